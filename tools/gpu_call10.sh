#!/bin/bash
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -o $O/${TAG}_full_$name -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_$name.log 2>&1
  ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_full_${name}_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_full_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_full_${name}_source.csv.gz
  rm -f $O/${TAG}_full_$name.ncu-rep
}
cap rbstep32 'tc_rbstep_kernel<\(int\)32' 27 9
cap rbstep64 'tc_rbstep_kernel<\(int\)64' 27 9
du -sh $O
