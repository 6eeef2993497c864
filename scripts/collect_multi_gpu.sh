#!/bin/bash
# Multi-GPU artefacts (run under `gpurun --gpus N` from the repo root): bitwise check of the row-sharded training against
# one GPU, bench lines of the sharded workloads. $1 = N, $2 = prefix, $3... = which parts (check ml20m netflix rank powerlaw1b)
N=$1; P=${2:-r02}; shift 2
O=gpurun_out
export WMF_SYNTH_CACHE=$PWD/.synth_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
set -x
for part in "$@"; do
  case $part in
    check)      timeout 240 $TR scripts/multi_gpu_check.py > $O/${P}_multi_gpu_${N}xB200.log 2>&1 ;;
    ml20m)      timeout 240 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${P}_bench_n${N}.log 2>&1 ;;
    netflix)    timeout 240 $TR bench.py --gpus $N --workload netflix --steps 5 --warmup 3 --no-e2e > $O/${P}_bench_n${N}_netflix.log 2>&1 ;;
    rank)       timeout 200 $TR bench.py --gpus $N --workload rank --steps 5 --warmup 3 > $O/${P}_bench_n${N}_rank.log 2>&1 ;;
    powerlaw1b) timeout 170 $TR bench.py --gpus $N --workload powerlaw1b --steps 3 --warmup 2 --no-e2e > $O/${P}_bench_n${N}_powerlaw1b.log 2>&1 ;;
  esac
  echo "$part rc=$?"
done
tail -c 600 $O/${P}_multi_gpu_${N}xB200.log 2>/dev/null
