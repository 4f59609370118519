#!/usr/bin/env bash
# First GPU call of round 2: everything the second (GPU-less) session of round 1 left to run, in priority order, each step
# under its own timeout so that one failure cannot eat the budget.  Results land in gpurun_out/ (merged back by gpurun).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round2_gpu_checklist.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {   # run <seconds> <logfile> <command...>
    local t=$1 log=$2; shift 2
    echo "== $* (timeout ${t}s)" | tee -a gpurun_out/checklist.log
    timeout "$t" "$@" > "gpurun_out/$log" 2>&1
    echo "   rc=$? $(tail -n 1 "gpurun_out/$log" | cut -c1-200)" | tee -a gpurun_out/checklist.log
}
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
# 1. the whole GPU suite: first hardware run of eval / GAN / full step / trainer / regraph / shard / large sgemm / SpMM prefetch variant
run 1500 pytest_gpu.log python -m pytest tests -m gpu -q -x --durations=15
# 1b. if -x stopped early, still get the verdict of every new file on its own
for f in tests/test_gpu_zz_eval.py tests/test_gpu_zz_gan.py tests/test_gpu_zz_fullstep.py tests/test_gpu_zz_more_ops.py tests/test_gpu_zz_trainer.py; do
    run 600 "pytest_$(basename "$f" .py).log" python -m pytest "$f" -m gpu -q
done
# 2. the tcgen05 wide GEMM and the CUDA-graph capture of the full step (gated: never run on hardware before)
run 600 pytest_gemm_wide.log env MMSSL_RUN_UNVALIDATED=1 python -m pytest tests/test_gpu_zzz_gemm_wide.py -m gpu -q
# 3. headline bench (unchanged path) + the stock-torch comparator
run 600 bench_default.json python bench.py
run 600 bench_sorted_items.json env MMSSL_SPMM_SORT=1 python bench.py --no-cpu-baseline
run 600 bench_stock_gpu.json python bench.py --impl stock-gpu --steps 100
# 4. SpMM probes incl. the early-prefetch variant (gcn_*_impl{4,68,16,80}) and the length-sorted work items (*_sorted_us)
run 600 probe_baby.json python tools/probe.py baby
run 600 probe_sports.json python tools/probe.py sports
# 5. the full training iteration: CUDA-core route, library route, tensor-core route
run 900 fullstep_simt.json python tools/fullstep_bench.py baby --gemm simt --steps 20 --cpu-steps 1 --phases
run 600 fullstep_cublas.json python tools/fullstep_bench.py baby --gemm cublas --steps 20 --cpu-steps 0
run 600 fullstep_tc.json python tools/fullstep_bench.py baby --gemm tc --steps 20 --cpu-steps 0 --phases
# 6. launch list of the full step (one iteration under ncu, serialised)
run 900 ncu_fullstep.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/fullstep_launches.csv \
    python tools/fullstep_bench.py baby --gemm simt --steps 1 --warmup 3 --cpu-steps 0
# 7. one full capture of the wide tensor-core GEMM and of the large CUDA-core GEMM (read here with `ncu -i ... --page raw --csv`)
run 900 ncu_gemm_wide.log ncu --set full --clock-control none --import-source on -k regex:gemm_wide_kernel -c 2 -f -o gpurun_out/gemm_wide \
    python tools/fullstep_bench.py baby --gemm tc --steps 1 --warmup 2 --cpu-steps 0
run 900 ncu_sgemm_large.log ncu --set full --clock-control none --import-source on -k regex:sgemm_large_kernel -c 2 -f -o gpurun_out/sgemm_large \
    python tools/fullstep_bench.py baby --gemm simt --steps 1 --warmup 2 --cpu-steps 0
# (multi-GPU, separate call with --gpus 2):  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/rowshard_step_bench.py sports check [mc]
echo done | tee -a gpurun_out/checklist.log
