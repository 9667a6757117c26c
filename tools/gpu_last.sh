#!/bin/bash
# last minutes of the round: the new defaults (lean issue, two-sub-tile 128-column stage) on the HiFi-GAN parity tests, one bench
# line each, and the A-once tiling of the block kernel again now that the issue loop no longer paces it
TAG=${1:-r2q}
O=gpurun_out
mkdir -p $O
( time timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hifigan and not subprocess and not loudness and not full_size" ) > $O/${TAG}_pytest_hifigan_defaults.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_hifigan_defaults.log; tail -3 $O/${TAG}_pytest_hifigan_defaults.log | cut -c1-200
timeout 40 python bench.py --workload hifigan --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
timeout 50 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn.json 2> $O/${TAG}_bench_pwn.err
CUBE_TC_AONCE=1 timeout 50 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_aonce.json 2> $O/${TAG}_bench_pwn_aonce.err
python - <<PY
import json
for f in ["hifigan", "pwn", "pwn_aonce"]:
    try:
        d = json.load(open("$O/${TAG}_bench_%s.json" % f))
        print(f, round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e6, 2), "M/s", d["clocks"]["sm_mhz"], "MHz", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
PY
