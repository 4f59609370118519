#!/usr/bin/env bash
# Round 2, SpMM call: GPU tests of the staged-gather SpMM + probes (small: in-graph latency; large: cold throughput).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-c5}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
run 300 ${TAG}_pytest_spmm.log python -m pytest tests/test_gpu_spmm_bulk.py -m gpu -q
run 300 ${TAG}_probe_small.json python tools/probe.py small
run 400 ${TAG}_probe_large.json python tools/probe.py large
echo done | tee -a gpurun_out/${TAG}.log
