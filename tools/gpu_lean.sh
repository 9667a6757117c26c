#!/bin/bash
# A/B + parity session for the lean issue loops (CUBE_TC_LEAN): benches first (is it faster at all?), then the GPU suite with the
# switch exported (= the suite as it will run once the default flips), then the instrumented block kernel and a launch list
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e6, 2), "M/s", d["clocks"]["sm_mhz"], "MHz", {k: round(v, 2) for k, v in d["layer_ms_last_step"].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
( time CUBE_TC_LEAN=1 timeout 90 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${TAG}_smoke_lean.log 2>&1
rc=$?; tail -4 $O/${TAG}_smoke_lean.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "smoke with CUBE_TC_LEAN=1 failed (rc=$rc): stopping"; exit 1; fi
for l in 0 1; do
  CUBE_TC_LEAN=$l timeout 90 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pwn_lean$l.json 2> $O/${TAG}_bench_pwn_lean$l.err
  show $O/${TAG}_bench_pwn_lean$l.json "student lean=$l"
  CUBE_TC_LEAN=$l timeout 60 python bench.py --workload hifigan --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_hifigan_lean$l.json 2> $O/${TAG}_bench_hifigan_lean$l.err
  show $O/${TAG}_bench_hifigan_lean$l.json "hifigan lean=$l"
done
CUBE_TC_LEAN=1 CUBE_TC_WIDE=1 timeout 100 python bench.py --workload hifigan --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_hifigan_lean1_wide1.json 2> $O/${TAG}_bench_hifigan_lean1_wide1.err
show $O/${TAG}_bench_hifigan_lean1_wide1.json "hifigan lean=1 wide=1"
( time CUBE_TC_LEAN=1 timeout 330 python -m pytest tests -m gpu -q -x -k "not (student_unfused_pair or student_cta_pair or hifigan_no_window or packed)" ) > $O/${TAG}_pytest_lean.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_lean.log
tail -6 $O/${TAG}_pytest_lean.log | cut -c1-300
CUBE_TC_LEAN=1 CUBE_BLOCK_STATS=1 timeout 100 python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_bench_pwn_lean1_stats.json 2> $O/${TAG}_block_stats_lean.txt
head -30 $O/${TAG}_block_stats_lean.txt
CUBE_TC_LEAN=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_pwn_lean.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_pwn.log 2>&1
CUBE_TC_LEAN=1 timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_block_kernel" -s 20 -c 1 -o $O/${TAG}_full_block_lean -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_full_block.log 2>&1
ncu -i $O/${TAG}_full_block_lean.ncu-rep --page raw --csv > $O/${TAG}_full_block_lean_raw.csv 2>/dev/null
ncu -i $O/${TAG}_full_block_lean.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_full_block_lean_source.csv.gz
rm -f $O/${TAG}_full_block_lean.ncu-rep
du -sh $O
