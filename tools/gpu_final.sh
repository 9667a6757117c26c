#!/bin/bash
# final evidence session of a round: tests, smoke, both bench arms with the driver's flags (timed), ncu of the final kernels
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
( time timeout 600 python bench.py --workload hifigan --impl reference --steps 10 --warmup 3 ) > $O/${TAG}_bench_hifigan_reference.json 2> $O/${TAG}_bench_hifigan_reference.err
( time timeout 600 python bench.py --workload hifigan --steps 10 --warmup 3 ) > $O/${TAG}_bench_hifigan.json 2> $O/${TAG}_bench_hifigan.err
CUBE_TC_CG2=1 timeout 300 python bench.py --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_hifigan_cg2.json 2> $O/${TAG}_bench_hifigan_cg2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_pwn.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_pwn.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_hifigan.csv python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hifigan.log 2>&1
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -o $O/${TAG}_full_$name -f python bench.py --workload hifigan --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_$name.log 2>&1
  ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_full_${name}_raw.csv 2>/dev/null
  rm -f $O/${TAG}_full_$name.ncu-rep
}
cap rbstep32 'tc_rbstep_kernel<\(int\)32' 27 9
cap rbstep64 'tc_rbstep_kernel<\(int\)64' 27 9
grep -h real $O/${TAG}_bench_*.err | head; du -sh $O
