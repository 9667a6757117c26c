#!/bin/bash
# ncu --set full of the wide-stage convs (k = 7 step of the 128- and 256-channel stages), summarised on the box
TAG=${1:-r2o}
O=gpurun_out
mkdir -p $O
cap() {  # name regex skip env
  env $4 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 2 -o $O/${TAG}_full_$1 -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_$1.log 2>&1
  ncu -i $O/${TAG}_full_$1.ncu-rep --page raw --csv > $O/${TAG}_full_$1_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_full_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_full_$1_source.csv.gz
  rm -f $O/${TAG}_full_$1.ncu-rep
  tail -2 $O/${TAG}_ncu_full_$1.log | cut -c1-200
}
cap conv128 'tc_conv_kernel<\(int\)128' 26 CUBE_TC_WIDE=0
cap conv256 'tc_conv_kernel<\(int\)256' 28 CUBE_TC_WIDE=0
cap conv128w 'tc_conv_kernel<\(int\)128' 62 CUBE_TC_WIDE=1
ls -la $O | grep ${TAG}; du -sh $O
