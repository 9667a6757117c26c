#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
( time timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
( time timeout 400 python bench.py --workload hifigan --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_hifigan_reference.json 2> $O/${TAG}_bench_hifigan_reference.err
CUBE_TC_PREFETCH=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "student and tcgen05 and not subprocess and not 862" > $O/${TAG}_pytest_prefetch.log 2>&1
echo "prefetch rc=$?" >> $O/${TAG}_pytest_prefetch.log; tail -2 $O/${TAG}_pytest_prefetch.log
CUBE_TC_PREFETCH=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_prefetch.json 2> $O/${TAG}_bench_pwn_prefetch.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_noprefetch.json 2> $O/${TAG}_bench_pwn_noprefetch.err
CUBE_TC_PREFETCH=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_prefetch2.json 2> $O/${TAG}_bench_pwn_prefetch2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], round(d["ms_per_step"],2), round(d["value"]/1e6,3), (d.get("clocks") or {}).get("sm_mhz"), d.get("wall_s"))
    except Exception as e: print(f, "ERR", e)
PY
grep -h real $O/${TAG}_bench_*.err
