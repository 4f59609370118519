#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-g2b}
N=${2:-2}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-700)" | tee -a gpurun_out/${TAG}.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run 300 ${TAG}_pytest_dist.log python -m pytest tests/test_gpu_dist.py tests/test_gpu_zz_dist_step.py -m gpu -q
run 200 ${TAG}_rs_syn1m_1stream.json env MMSSL_ROWSHARD_STREAMS=0 $TR --master-port 29801 tools/rowshard_step_bench.py syn1m check mc graph --steps 20
run 200 ${TAG}_rs_syn1m_2stream.json $TR --master-port 29802 tools/rowshard_step_bench.py syn1m check mc graph --steps 20
run 200 ${TAG}_rs_sports_2stream.json $TR --master-port 29803 tools/rowshard_step_bench.py sports check mc graph --steps 50
run 400 ${TAG}_bench.json $TR --master-port 29805 bench.py --gpus $N --steps 200 --warmup 10 --row-shard sports
echo done | tee -a gpurun_out/${TAG}.log
