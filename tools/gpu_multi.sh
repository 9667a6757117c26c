#!/bin/bash
# usage: bash tools/gpu_multi.sh <N> <tag>
N=${1:-2}
TAG=${2:-r2n$N}
O=gpurun_out
mkdir -p $O
run() {  # name args...
  local name=$1; shift
  ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" ) > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err
  tail -c 600 $O/${TAG}_bench_${name}.json | head -c 400; echo
}
run pwn --steps 10 --warmup 3 --no-cpu-baseline
run ragged --workload ragged --steps 8 --warmup 3
run e2e --workload e2e --steps 4 --warmup 3
run hifigan --workload hifigan --steps 10 --warmup 3 --no-cpu-baseline
grep -h "NCCL INFO" $O/${TAG}_bench_ragged.err | grep -i "nvls\|p2p\|channel\|via" | head -12 > $O/${TAG}_nccl.txt
du -sh $O
