#!/usr/bin/env bash
# Round 2: BASELINE config 5 (10M x 1M, 200M edges, d = 256) row-sharded on the GPUs of one box; a scaled-down run first.
#   gpurun --gpus N -- bash tools/r2_call_c5.sh TAG N [small|full|both]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-c5}; N=${2:-8}; WHAT=${3:-both}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-400)" | tee -a gpurun_out/${TAG}.log; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
if [ "$WHAT" != full ]; then
  run 300 ${TAG}_small.json $T tools/config5_run.py --users 400000 --items 80000 --edges 6000000 --embed-size 256 --dv 512 --dt 256 --steps 5
fi
if [ "$WHAT" != small ]; then
  run 600 ${TAG}_full.json $T tools/config5_run.py --steps 5
fi
echo done | tee -a gpurun_out/${TAG}.log
