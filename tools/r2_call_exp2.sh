#!/usr/bin/env bash
# Round 2: where the SpMM plan cuts rows; 24 KB id-fusion kernels beside the GEMMs; wider dWcat reduce.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x2}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
run 300 ${TAG}_pytest.log python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_more_ops.py tests/test_gpu_model.py -m gpu -q -x
run 120 ${TAG}_probe_kernels.json python tools/probe.py kernels
run 300 ${TAG}_probe_plan.json python tools/probe.py plan
B="python bench.py --extra-configs none --no-cpu-baseline"
run 200 ${TAG}_bench_default.json $B
run 200 ${TAG}_trace_default.txt python tools/trace_step.py baby
run 200 ${TAG}_bench_cuts32_16.json env MMSSL_SPMM_CUTS=32,16,1024,64 $B
run 200 ${TAG}_bench_cuts32_32.json env MMSSL_SPMM_CUTS=32,32,1024,64 $B
run 200 ${TAG}_bench_cuts16_16.json env MMSSL_SPMM_CUTS=16,16,512,32 $B
run 200 ${TAG}_bench_cuts24_12.json env MMSSL_SPMM_CUTS=24,12,512,32 $B
run 200 ${TAG}_trace_cuts32_16.txt env MMSSL_SPMM_CUTS=32,16,1024,64 python tools/trace_step.py baby
echo done | tee -a gpurun_out/${TAG}.log
