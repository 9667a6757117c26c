#!/bin/bash
# keeps gpurun_out/ small: full captures are summarised ON the box (raw + source CSV, gzip) and the .ncu-rep deleted
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
CUBE_BLOCK_STATS=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_stats.json 2> $O/${TAG}_block_stats.txt
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn.json 2> $O/${TAG}_bench_pwn.err
( time timeout 400 python bench.py --workload hifigan --steps 10 --warmup 3 ) > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
CUBE_TC_RBFUSE=0 timeout 200 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan_unfused.json 2> $O/${TAG}_bench_hifigan_unfused.err
timeout 200 python bench.py --workload api1 --steps 20 --warmup 3 > $O/${TAG}_bench_api1.json 2> $O/${TAG}_bench_api1.err
CUBE_GRAPH=0 timeout 200 python bench.py --workload api1 --steps 20 --warmup 3 > $O/${TAG}_bench_api1_nograph.json 2> $O/${TAG}_bench_api1_nograph.err
timeout 200 python bench.py --workload e2e --steps 5 --warmup 3 > $O/${TAG}_bench_e2e_n1.json 2> $O/${TAG}_bench_e2e_n1.err
timeout 200 python bench.py --workload ragged --steps 5 --warmup 3 > $O/${TAG}_bench_ragged_n1.json 2> $O/${TAG}_bench_ragged_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
cap() {  # name regex skip env
  local name=$1 rx=$2 skip=$3
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 2 -o $O/${TAG}_full_$name -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_$name.log 2>&1
  ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_full_${name}_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_full_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_full_${name}_source.csv.gz
  rm -f $O/${TAG}_full_$name.ncu-rep
}
cap rbstep32 'tc_rbstep_kernel<\(int\)32' 12
cap rbstep64 'tc_rbstep_kernel<\(int\)64' 12
export CUBE_TC_RBFUSE=0
cap conv32 'tc_conv_kernel<\(int\)32' 40
cap conv64 'tc_conv_kernel<\(int\)64' 40
unset CUBE_TC_RBFUSE
SEL="hifigan_golden_mini or hifigan_config_v1_random_weights_ragged or student_small_random_weights or upsample2_golden or mulaw_bit_exact or wavernn_golden or mel_cube_flavour"
( time timeout 420 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" ) > $O/${TAG}_sanitizer_memcheck.log 2>&1
( time timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hifigan_config_v1_random_weights_ragged or student_small_random_weights" ) > $O/${TAG}_sanitizer_racecheck.log 2>&1
tail -n 6 $O/${TAG}_sanitizer_memcheck.log; tail -n 6 $O/${TAG}_sanitizer_racecheck.log
du -sh $O; ls -la $O | grep $TAG | awk '{print $5, $9}'
