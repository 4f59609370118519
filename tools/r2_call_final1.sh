#!/usr/bin/env bash
# Round 2, final 1-GPU call: whole GPU suite, compute-sanitizer passes, the default bench line, traces, ncu of the changed kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-f1}
run() { local t=$1 log=$2; shift 2; echo "== $* (timeout ${t}s)" | tee -a gpurun_out/${TAG}.log
        local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "   rc=$? $((SECONDS-t0))s $(tail -n 1 "gpurun_out/$log" | cut -c1-300)" | tee -a gpurun_out/${TAG}.log; }
run 600 ${TAG}_pytest_gpu.log python -m pytest tests -m gpu -q
run 500 ${TAG}_bench.json python bench.py
run 200 ${TAG}_trace_baby.txt python tools/trace_step.py baby
run 200 ${TAG}_trace_sports.txt python tools/trace_step.py sports
run 120 ${TAG}_probe_kernels.json python tools/probe.py kernels
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="tests/test_gpu_ops.py::test_spmm_plain[1-64] tests/test_gpu_ops.py::test_spmm_plain[2-128] tests/test_gpu_ops.py::test_spmm_plain[3-256] tests/test_gpu_ops.py::test_spmm_epilogues[64] tests/test_gpu_spmm_bulk.py::test_spmm_bulk_plain[variant1-1-64] tests/test_gpu_spmm_bulk.py::test_spmm_bulk_plain[variant3-2-128] tests/test_gpu_spmm_bulk.py::test_spmm_bulk_heavy_rows_and_zipf_columns tests/test_gpu_ops.py::test_infonce_forward_backward[257-64] tests/test_gpu_ops.py::test_infonce_forward_backward[1024-64] tests/test_gpu_zz_more_ops.py::test_spmm_pipelined_walk_matches_default[0-64-1-7] tests/test_gpu_zz_more_ops.py::test_spmm_pipelined_walk_matches_default[64-128-2-11] tests/test_gpu_zz_more_ops.py::test_spmm_plan_cuts[64-1-cuts0]"
for tool in racecheck memcheck synccheck; do
  run 300 ${TAG}_${tool}.log $CS --tool $tool --print-limit 10 python -m pytest $SEL -m gpu -q -v
done
NCU="ncu --set full --clock-control none -f"
run 300 ${TAG}_ncu_step.log $NCU -k regex:"nce_stats_tc_kernel|id_fuse2_fwd_kernel|id_fuse2_bwd_kernel|dwcat_reduce_kernel" -s 8 -c 6 -o gpurun_out/${TAG}_step_baby python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-configs none
[ -f gpurun_out/${TAG}_step_baby.ncu-rep ] && ncu -i gpurun_out/${TAG}_step_baby.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_baby_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out | tee -a gpurun_out/${TAG}.log
echo done | tee -a gpurun_out/${TAG}.log
