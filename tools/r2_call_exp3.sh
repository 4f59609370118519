#!/usr/bin/env bash
# Round 2: shared-memory carve-out preferences (kernels that should co-reside with the projection GEMMs' CTAs).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x3}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
B="python bench.py --extra-configs none --no-cpu-baseline"
run 200 ${TAG}_bench_default.json $B
run 200 ${TAG}_trace_default.txt python tools/trace_step.py baby
run 200 ${TAG}_bench_spmm100.json env MMSSL_SPMM_CARVEOUT=100 $B
run 200 ${TAG}_trace_spmm100.txt env MMSSL_SPMM_CARVEOUT=100 python tools/trace_step.py baby
run 200 ${TAG}_bench_spmm50.json env MMSSL_SPMM_CARVEOUT=50 $B
run 200 ${TAG}_bench_idfuse_default_carveout.json env MMSSL_IDFUSE_CARVEOUT=-1 $B
run 120 ${TAG}_probe_kernels.json python tools/probe.py kernels
echo done | tee -a gpurun_out/${TAG}.log
