#!/bin/bash
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "hifigan and not subprocess" ) > $O/${TAG}_pytest_hifigan.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_hifigan.log
tail -4 $O/${TAG}_pytest_hifigan.log
timeout 300 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -o $O/${TAG}_full_$name -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_$name.log 2>&1
  ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_full_${name}_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_full_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_full_${name}_source.csv.gz
  rm -f $O/${TAG}_full_$name.ncu-rep
}
# launches 27..35 of rbstep<32> in the 4th forward = ResBlocks k=3 (3 steps), k=7 (3), k=11 (3) of the last stage
cap rbstep32 'tc_rbstep_kernel<\(int\)32' 27 9
cap rbstep64 'tc_rbstep_kernel<\(int\)64' 27 9
grep -h "ms_per_step" $O/${TAG}_bench_hifigan.json | cut -c1-300
du -sh $O
