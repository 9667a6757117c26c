#!/bin/bash
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log | cut -c1-300
( time timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
( time timeout 300 python bench.py --workload hifigan --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_hifigan_reference.json 2> $O/${TAG}_bench_hifigan_reference.err
( time timeout 300 python bench.py --workload hifigan --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
timeout 120 python bench.py --workload ragged --steps 5 --warmup 3 > $O/${TAG}_bench_ragged_n1.json 2> $O/${TAG}_bench_ragged_n1.err
timeout 120 python bench.py --workload api1 --steps 20 --warmup 3 > $O/${TAG}_bench_api1.json 2> $O/${TAG}_bench_api1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
grep -h real $O/${TAG}_bench_*.err | tr '\n' ' '; du -sh $O
