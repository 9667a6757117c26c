#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn.json 2> $O/${TAG}_bench_pwn.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
python - <<'PY'
import json
for f in ("hifigan","pwn"):
    d=json.load(open("gpurun_out/%s_bench_%s.json" % ("r2e", f)))
    print(f, round(d["ms_per_step"],2), d["layer_ms_last_step"], d["clocks"])
PY
du -sh $O
