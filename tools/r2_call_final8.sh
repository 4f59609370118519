#!/usr/bin/env bash
# Round 2, final multi-GPU call: bench.py --gpus N (data-parallel headline + parity + row-sharded Sports / 1M x 200k).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-f8}
N=${2:-8}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-600)" | tee -a gpurun_out/${TAG}.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
run 500 ${TAG}_bench.json $TR --master-port 29705 bench.py --gpus $N --steps 200 --warmup 10
echo done | tee -a gpurun_out/${TAG}.log
