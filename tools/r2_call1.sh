#!/usr/bin/env bash
# Round 2, GPU call 1 (1 GPU): full GPU suite, wide-GEMM accumulation sweep, headline bench + stock-torch comparator, SpMM probes,
# full-step timings.  Each step under its own timeout; results in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/call1.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/call1.log; }
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
run 400 c1_pytest_gpu.log python -m pytest tests -m gpu -q --durations=10
run 200 c1_gemm_wide_sweep.json python tools/gemm_wide_sweep.py
run 300 c1_pytest_gemm_wide_unvalidated.log env MMSSL_RUN_UNVALIDATED=1 python -m pytest tests/test_gpu_zzz_gemm_wide.py -m gpu -q
run 300 c1_bench_default.json python bench.py
run 200 c1_bench_stock_baby.json python bench.py --impl stock-gpu --steps 100
run 200 c1_bench_stock_sports.json python bench.py --impl stock-gpu --steps 100 --config sports
run 200 c1_probe_small.json python tools/probe.py small
run 300 c1_probe_large.json python tools/probe.py large
run 300 c1_fullstep_simt.json python tools/fullstep_bench.py baby --gemm simt --steps 20 --cpu-steps 0 --phases
run 200 c1_fullstep_cublas.json python tools/fullstep_bench.py baby --gemm cublas --steps 20 --cpu-steps 0
run 200 c1_fullstep_tc.json python tools/fullstep_bench.py baby --gemm tc --steps 20 --cpu-steps 0 --phases
echo done | tee -a gpurun_out/call1.log
