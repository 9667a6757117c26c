#!/bin/bash
# one GPU-box session: tests, both bench arms with the driver's flags, launch lists and full ncu captures.
# usage (from the repo root, on the box): bash tools/gpu_call.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/${TAG}_smi.txt 2>&1
nproc > $O/${TAG}_host.txt; cat /sys/fs/cgroup/cpu.max >> $O/${TAG}_host.txt 2>&1; free -g >> $O/${TAG}_host.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
( time timeout 600 python bench.py --workload hifigan --impl reference --steps 10 --warmup 3 ) > $O/${TAG}_bench_hifigan_reference.json 2> $O/${TAG}_bench_hifigan_reference.err
( time timeout 600 python bench.py --workload hifigan --steps 10 --warmup 3 ) > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
( time timeout 600 python bench.py --workload ragged --steps 5 --warmup 3 ) > $O/${TAG}_bench_ragged_n1.json 2> $O/${TAG}_bench_ragged_n1.err
# launch lists (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_pwn.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_pwn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tc_block_pair_kernel -s 20 -c 1 -o $O/${TAG}_full_pair -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_conv_kernel<32," -s 40 -c 2 -o $O/${TAG}_full_conv32 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_conv32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_conv_kernel<64," -s 40 -c 2 -o $O/${TAG}_full_conv64 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_conv64.log 2>&1
ls -la $O | tail -30
