#!/bin/bash
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -8 $O/${TAG}_pytest.log
# student: relay vs direct peer->leader signalling
CUBE_PAIR_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "student and tcgen05 and not 862 and not subprocess" > $O/${TAG}_pytest_direct.log 2>&1
echo "direct rc=$?" >> $O/${TAG}_pytest_direct.log; tail -3 $O/${TAG}_pytest_direct.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_relay.json 2> $O/${TAG}_bench_pwn_relay.err
CUBE_PAIR_DIRECT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_direct.json 2> $O/${TAG}_bench_pwn_direct.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_relay2.json 2> $O/${TAG}_bench_pwn_relay2.err
# hifigan: fused ResBlock steps (default) vs the conv pair
( time timeout 600 python bench.py --workload hifigan --steps 10 --warmup 3 ) > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
CUBE_TC_RBFUSE=0 timeout 300 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan_unfused.json 2> $O/${TAG}_bench_hifigan_unfused.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_rbstep_kernel<\(int\)32" -s 12 -c 2 -o $O/${TAG}_full_rbstep32 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_rbstep32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_rbstep_kernel<\(int\)64" -s 12 -c 2 -o $O/${TAG}_full_rbstep64 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_rbstep64.log 2>&1
CUBE_TC_RBFUSE=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_conv_kernel<\(int\)32" -s 40 -c 2 -o $O/${TAG}_full_conv32 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_conv32.log 2>&1
CUBE_TC_RBFUSE=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_conv_kernel<\(int\)64" -s 40 -c 2 -o $O/${TAG}_full_conv64 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_conv64.log 2>&1
# sanitizer on the small parity tests of every kernel family
SEL="hifigan_golden_mini or hifigan_config_v1_random_weights_ragged or student_small_random_weights or upsample2_golden or mulaw_bit_exact or wavernn_golden or mel_cube_flavour"
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" ) > $O/${TAG}_sanitizer_memcheck.log 2>&1
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hifigan_config_v1_random_weights_ragged or student_small_random_weights" ) > $O/${TAG}_sanitizer_racecheck.log 2>&1
tail -5 $O/${TAG}_sanitizer_memcheck.log $O/${TAG}_sanitizer_racecheck.log
ls -la $O | grep $TAG
