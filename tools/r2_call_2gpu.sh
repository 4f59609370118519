#!/usr/bin/env bash
# Round 2, 2-GPU call: multi-GPU parity tests (skipped on the 1-GPU test box), the row-sharded whole hot step, bench.py --gpus 2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/call_2gpu.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-400)" | tee -a gpurun_out/call_2gpu.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/gpus_2.txt 2>&1
run 400 g2_pytest_dist.log python -m pytest tests/test_gpu_dist.py tests/test_gpu_zz_dist_step.py -m gpu -q
run 300 g2_rowshard_sports_mc.json $TR --master-port 29601 tools/rowshard_step_bench.py sports check mc --steps 50
run 300 g2_rowshard_sports_mc_graph.json $TR --master-port 29602 tools/rowshard_step_bench.py sports check mc graph --steps 50
run 400 g2_bench.json $TR --master-port 29603 bench.py --gpus 2 --steps 200 --warmup 10
echo done | tee -a gpurun_out/call_2gpu.log
