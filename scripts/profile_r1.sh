#!/bin/bash
# Round-1 profile recipe (run under gpurun, one GPU): plain bench first, then the ncu passes.
set -x
ARGS="--steps 1 --warmup 1 --freqs-per-step 1 --streams 1 --no-cpu-baseline"
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py $ARGS > gpurun_out/bench_prof_plain.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1
# first look-ahead trailing update of a factorisation: zgemm3m launches 0-6 are the block-column updates of
# outer block 0, 7 is T1, 8 is T2
ncu --set full --clock-control none --import-source on -k regex:zgemm3m --launch-skip 8 --launch-count 1 \
    -o gpurun_out/prof_zgemm3m_r1c -f python bench.py $ARGS > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_regular_sym --launch-count 1 \
    -o gpurun_out/prof_regular_sym_r1c -f python bench.py $ARGS > gpurun_out/ncu_full2.log 2>&1
# batched H-matvec kernels (8x8 array, 64 right-hand sides): t = V^T x and y|band
ncu --set full --clock-control none --import-source on -k regex:k_hmv_gemm --launch-skip 2 --launch-count 2 \
    -o gpurun_out/prof_hmv_gemm_r1c -f python scripts/time_hmv.py 64 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/
