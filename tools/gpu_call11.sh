#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "hifigan and not subprocess" > $O/${TAG}_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
if [ $rc -eq 0 ]; then
  timeout 120 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
  python -c "
import json; d=json.load(open('$O/${TAG}_bench_hifigan.json')); print(round(d['ms_per_step'],2), round(d['value']/1e6,2), d['clocks']['sm_mhz'], d['layer_ms_last_step'])"
fi
