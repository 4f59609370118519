#!/usr/bin/env bash
# Round 2: first hardware run of the tensor-core InfoNCE + the whole suite + bench + kernel probes + trace.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-n1}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-400)" | tee -a gpurun_out/${TAG}.log; }
run 400 ${TAG}_pytest_gpu.log python -m pytest tests -m gpu -q
run 200 ${TAG}_probe_kernels.json python tools/probe.py kernels
run 400 ${TAG}_bench.json python bench.py --extra-configs none
run 200 ${TAG}_trace_baby.txt python tools/trace_step.py baby
echo done | tee -a gpurun_out/${TAG}.log
