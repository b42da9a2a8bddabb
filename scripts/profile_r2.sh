#!/bin/bash
# Round-2 profile recipe (run under gpurun, one GPU): plain runs first, then the ncu passes.
set -x
ARGS="--steps 1 --warmup 1 --freqs-per-step 1 --streams 1 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/r2_bench_prof_plain.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py $ARGS > gpurun_out/r2_ncu_launch.log 2>&1
# first look-ahead trailing update of a factorisation (zgemm3m launches 0-6: block-column updates of outer block 0, 7: T1, 8: T2)
ncu --set full --clock-control none --import-source on -k regex:zgemm3m --launch-skip 8 --launch-count 1 \
    -o gpurun_out/r2_prof_zgemm3m_T2 -f python bench.py $ARGS > gpurun_out/r2_ncu_full1.log 2>&1
# many-source field pressure: node-space pressure vectors + contraction (C3 mesh, one chunk of observation points)
python scripts/time_pfield.py 75776 3 > gpurun_out/r2_pfield_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_pvec_tiles --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2_prof_pvec -f python scripts/time_pfield.py 75776 3 > gpurun_out/r2_ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:zgemm3m --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2_prof_zgemm3m_pfield -f python scripts/time_pfield.py 75776 3 > gpurun_out/r2_ncu_full3.log 2>&1
# batched H-matvec (8x8 array, 96 right-hand sides): stacked t-pass and y-pass
python scripts/time_hmv.py 8 15 5 96 > gpurun_out/r2_hmv_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_hmv_tstack --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2_prof_hmv_tstack -f python scripts/time_hmv.py 8 15 5 96 > gpurun_out/r2_ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hmv_gemm --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2_prof_hmv_ypass -f python scripts/time_hmv.py 8 15 5 96 > gpurun_out/r2_ncu_full5.log 2>&1
# time-domain FIR convolution (P = 144, 4000 taps)
python scripts/time_firconv.py > gpurun_out/r2_firconv_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_fir_conv --launch-skip 2 --launch-count 1 \
    -o gpurun_out/r2_prof_firconv -f python scripts/time_firconv.py > gpurun_out/r2_ncu_full6.log 2>&1
ls -la gpurun_out/ | grep r2_
