#!/usr/bin/env bash
# Round 2: one-matrix-per-CTA tensor-core InfoNCE, user-side fusion backward beside the chain, software-pipelined SpMM walk,
# stream priorities: parity tests, kernel probes, and the Baby bench line under each setting.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x1}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
run 300 ${TAG}_pytest.log python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_more_ops.py tests/test_gpu_model.py -m gpu -q -x
run 120 ${TAG}_probe_kernels.json python tools/probe.py kernels
run 200 ${TAG}_probe_pipe.json python tools/probe.py pipe
B="python bench.py --extra-configs none --no-cpu-baseline"
run 200 ${TAG}_bench_default.json $B
run 200 ${TAG}_trace_default.txt python tools/trace_step.py baby
run 200 ${TAG}_bench_side1.json env MMSSL_PRIO_SIDE=-1 $B
run 200 ${TAG}_bench_side2.json env MMSSL_PRIO_SIDE=-2 MMSSL_PRIO_FORK=-2 $B
run 200 ${TAG}_bench_pipe.json env MMSSL_SPMM_SMALL_IMPL=516 $B
run 200 ${TAG}_bench_pipe_pre.json env MMSSL_SPMM_SMALL_IMPL=580 $B
run 200 ${TAG}_bench_pipe_side1.json env MMSSL_SPMM_SMALL_IMPL=516 MMSSL_PRIO_SIDE=-1 $B
run 200 ${TAG}_trace_pipe.txt env MMSSL_SPMM_SMALL_IMPL=516 python tools/trace_step.py baby
echo done | tee -a gpurun_out/${TAG}.log
