#!/bin/bash
# A/B session for the opt-in HiFi-GAN variants (CUBE_TC_WIDE, CUBE_RB_PK): targeted parity tests, then the same bench on one box
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
( time timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "two_subtile or packed" ) > $O/${TAG}_pytest_variants.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_variants.log
tail -5 $O/${TAG}_pytest_variants.log | cut -c1-400
for v in "0 0" "1 0" "0 1" "1 1"; do
  set -- $v
  CUBE_TC_WIDE=$1 CUBE_RB_PK=$2 timeout 200 python bench.py --workload hifigan --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_hifigan_w$1_p$2.json 2> $O/${TAG}_bench_hifigan_w$1_p$2.err
  python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_hifigan_w$1_p$2.json"))
    print("wide=$1 pk=$2", round(d["ms_per_step"], 2), "ms", d["clocks"]["sm_mhz"], "MHz", {k: round(v, 2) for k, v in d["layer_ms_last_step"].items()})
except Exception as e:
    print("wide=$1 pk=$2 FAILED", e)
PY
done
CUBE_TC_WIDE=1 CUBE_RB_PK=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan_w1_p1.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
CUBE_TC_WIDE=1 CUBE_RB_PK=1 timeout 120 python bench.py --workload ragged --steps 5 --warmup 3 > $O/${TAG}_bench_ragged_n1_w1_p1.json 2> $O/${TAG}_bench_ragged_n1.err
CUBE_TC_WIDE=1 CUBE_RB_PK=1 timeout 120 python bench.py --workload api1 --steps 20 --warmup 3 > $O/${TAG}_bench_api1_w1_p1.json 2> $O/${TAG}_bench_api1.err
du -sh $O
