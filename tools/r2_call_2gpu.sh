#!/usr/bin/env bash
# Round 2, 2-GPU call: multi-GPU parity tests (skipped on the 1-GPU test box), the row-sharded whole hot step, bench.py --gpus 2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-g2}
N=${2:-2}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-600)" | tee -a gpurun_out/${TAG}.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
if [ "$N" = "2" ]; then run 400 ${TAG}_pytest_dist.log python -m pytest tests/test_gpu_dist.py tests/test_gpu_zz_dist_step.py -m gpu -q; fi
run 200 ${TAG}_rs_tiktok_nccl.json $TR --master-port 29601 tools/rowshard_step_bench.py tiktok check --steps 20
run 200 ${TAG}_rs_sports_mc.json $TR --master-port 29602 tools/rowshard_step_bench.py sports check mc --steps 50
run 200 ${TAG}_rs_sports_mc_ag.json $TR --master-port 29603 tools/rowshard_step_bench.py sports check mc allgather --steps 50
run 200 ${TAG}_rs_sports_mc_graph.json $TR --master-port 29604 tools/rowshard_step_bench.py sports check mc graph --steps 50
run 500 ${TAG}_bench.json $TR --master-port 29605 bench.py --gpus $N --steps 200 --warmup 10
echo done | tee -a gpurun_out/${TAG}.log
