#!/usr/bin/env bash
# Round 2: compute-sanitizer passes + ncu captures.  The .ncu-rep files are exported to CSV ON THE BOX and deleted: gpurun copies
# back at most 64 MiB of gpurun_out/ (a first version of this script lost its results to that limit).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p2}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
CS=/usr/local/cuda/bin/compute-sanitizer
# 1. sanitizer on the split-row SpMM (both kernels) and the tensor-core InfoNCE: a handful of tests each (the tools slow kernels 10-50x)
run 240 ${TAG}_racecheck.log $CS --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_spmm_bulk.py tests/test_gpu_ops.py -m gpu -q -k "test_spmm_bulk_plain and 64-1 or test_spmm_plain and 64-1 or test_spmm_bulk_heavy or test_infonce_forward_backward and 257"
run 240 ${TAG}_memcheck.log $CS --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_spmm_bulk.py tests/test_gpu_ops.py -m gpu -q -k "test_spmm_bulk_plain and 64-1 or test_spmm_plain and 64-1 or test_spmm_bulk_heavy or test_spmm_bulk_many or test_infonce_forward_backward and 257"
# 2. ncu --set full: the LDG SpMM at 1M x 200k without / with the L2 residency hints, the staged-gather kernel
export_rep() { for f in gpurun_out/${TAG}_$1.ncu-rep; do [ -f "$f" ] && ncu -i "$f" --page raw --csv > "gpurun_out/${TAG}_$1_raw.csv" 2>/dev/null; rm -f "$f"; done; }
NCU="ncu --set full --clock-control none -f"
run 300 ${TAG}_ncu_spmm_syn1m_ldg.log $NCU -k regex:spmm_csr_kernel -s 2 -c 1 -o gpurun_out/${TAG}_spmm_syn1m_ldg python tools/spmm_large_ncu.py 16
export_rep spmm_syn1m_ldg
run 300 ${TAG}_ncu_spmm_syn1m_hint.log $NCU -k regex:spmm_csr_kernel -s 2 -c 1 -o gpurun_out/${TAG}_spmm_syn1m_hint python tools/spmm_large_ncu.py 144
export_rep spmm_syn1m_hint
run 300 ${TAG}_ncu_spmm_syn1m_staged.log $NCU -k regex:spmm_bulk_kernel -s 2 -c 1 -o gpurun_out/${TAG}_spmm_syn1m_staged python tools/spmm_large_ncu.py 1048576
export_rep spmm_syn1m_staged
# 3. ncu: dominant kernels inside the step at Baby (one graph-free eager step is profiled: bench's launch counting step)
run 400 ${TAG}_ncu_step.log $NCU -k regex:"spmm_csr_kernel|gemm_bf16x3_kernel|nce_stats_tc_kernel" -s 30 -c 14 -o gpurun_out/${TAG}_step_baby python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-configs none
export_rep step_baby
# 4. launch list of the same command (cold-cache, serialised: compare shares)
run 400 ${TAG}_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-configs none
du -sh gpurun_out | tee -a gpurun_out/${TAG}.log
echo done | tee -a gpurun_out/${TAG}.log
