#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
CUBE_TC_AONCE=1 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "student and tcgen05 and not subprocess" > $O/${TAG}_pytest_aonce.log 2>&1
echo "aonce rc=$?" >> $O/${TAG}_pytest_aonce.log; tail -3 $O/${TAG}_pytest_aonce.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_default.json 2> $O/${TAG}_bench_pwn_default.err
CUBE_TC_AONCE=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_aonce.json 2> $O/${TAG}_bench_pwn_aonce.err
CUBE_TC_AONCE=1 CUBE_TC_FP8=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_aonce_fp16.json 2> $O/${TAG}_bench_pwn_aonce_fp16.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_default2.json 2> $O/${TAG}_bench_pwn_default2.err
timeout 200 python bench.py --workload ragged --steps 5 --warmup 3 > $O/${TAG}_bench_ragged_n1.json 2> $O/${TAG}_bench_ragged_n1.err
timeout 200 python bench.py --workload ragged --steps 8 --warmup 3 --emulate-world 8 > $O/${TAG}_bench_ragged_emu8.json 2> $O/${TAG}_bench_ragged_emu8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], round(d["ms_per_step"],2), round(d["value"]/1e6,2), (d.get("clocks") or {}).get("sm_mhz"), d["config"].get("batches"))
    except Exception as e: print(f, "ERR", e)
PY
