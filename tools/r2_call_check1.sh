#!/usr/bin/env bash
# Round 2: whole GPU suite + default bench line after the last kernel changes (tile height by table size, spill-free capped SpMM).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-k1}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
run 300 ${TAG}_pytest_gpu.log python -m pytest tests -m gpu -q -x
run 300 ${TAG}_bench.json python bench.py --no-cpu-baseline
echo done | tee -a gpurun_out/${TAG}.log
