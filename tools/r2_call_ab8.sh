#!/usr/bin/env bash
# Round 2: the row-sharded step with one vs two streams on N GPUs (1M x 200k), same process.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-ab8}; N=${2:-8}; shift 2 || true
CONFIGS=${*:-syn1m}
echo "== rowshard_ab $CONFIGS on $N GPUs" | tee -a gpurun_out/${TAG}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 tools/rowshard_ab.py $CONFIGS --steps 20 --tag ${TAG}_lines > gpurun_out/${TAG}_out.txt 2>&1
echo "   rc=$? $(tail -n 2 gpurun_out/${TAG}_out.txt | cut -c1-300)" | tee -a gpurun_out/${TAG}.log
echo done | tee -a gpurun_out/${TAG}.log
