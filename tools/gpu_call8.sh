#!/bin/bash
TAG=${1:-r2i}
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "(hifigan or cubegan or upsamplenet or student_small or forward_host) and not subprocess" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn.json 2> $O/${TAG}_bench_pwn.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], round(d["ms_per_step"],2), round(d["value"]/1e6,3), (d.get("clocks") or {}).get("sm_mhz"), d["layer_ms_last_step"])
    except Exception as e: print(f, "ERR", e)
PY
