#!/usr/bin/env bash
# Round 2, GPU call 2: first hardware run of the bulk-copy SpMM (tests + probes at baby / sports / syn1m).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/call3.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/call3.log; }
run 300 c3_pytest_bulk.log python -m pytest tests/test_gpu_spmm_bulk.py -m gpu -q -x
run 300 c3_probe_small.json python tools/probe.py small
run 400 c3_probe_large.json python tools/probe.py large
echo done | tee -a gpurun_out/call3.log
