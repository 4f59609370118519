#!/usr/bin/env bash
# Round 2: compute-sanitizer passes + ncu captures (read back in the build container with `ncu -i ... --page raw --csv`).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p1}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
CS=/usr/local/cuda/bin/compute-sanitizer
# 1. sanitizer: shared-memory races + sync misuse + out-of-bounds on the split-row SpMM (both kernels) and the tensor-core InfoNCE
run 900 ${TAG}_racecheck_spmm.log $CS --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_spmm_bulk.py tests/test_gpu_ops.py -m gpu -q -k "plain or epilogues or heavy or short or infonce"
run 900 ${TAG}_memcheck_spmm.log $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_spmm_bulk.py tests/test_gpu_ops.py -m gpu -q -k "plain or epilogues or heavy or short or infonce or empty"
run 600 ${TAG}_synccheck_spmm.log $CS --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_spmm_bulk.py -m gpu -q -k "plain or epilogues"
# 2. ncu: the LDG SpMM at 1M x 200k with and without the L2 residency hints, and the staged-gather kernel
NCU="ncu --set full --clock-control none --import-source on -f"
run 600 ${TAG}_ncu_spmm_syn1m_ldg.log $NCU -k regex:spmm_csr_kernel -s 2 -c 1 -o gpurun_out/${TAG}_spmm_syn1m_ldg python tools/spmm_large_ncu.py 16
run 600 ${TAG}_ncu_spmm_syn1m_hint.log $NCU -k regex:spmm_csr_kernel -s 2 -c 1 -o gpurun_out/${TAG}_spmm_syn1m_hint python tools/spmm_large_ncu.py 144
run 600 ${TAG}_ncu_spmm_syn1m_staged.log $NCU -k regex:spmm_bulk_kernel -s 2 -c 1 -o gpurun_out/${TAG}_spmm_syn1m_staged python tools/spmm_large_ncu.py 1048576
# 3. ncu: the dominant kernels inside the step at Baby (SpMM, projection GEMM, tensor-core InfoNCE)
run 900 ${TAG}_ncu_step.log $NCU -k regex:"spmm_csr_kernel|gemm_bf16x3_kernel|nce_stats_tc_kernel" -s 40 -c 24 -o gpurun_out/${TAG}_step_baby python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-configs none
# 4. launch list of the same command (cold-cache, serialised: compare shares)
run 900 ${TAG}_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-configs none
echo done | tee -a gpurun_out/${TAG}.log
