#!/bin/bash
# Final round-2 collection on one B200 (run under gpurun): GPU tests, the three bench configs, auxiliary timings,
# and the ncu captures of the time-domain kernels and the single-vector H-matvec -> gpurun_out/r2z_*
set -x
OUT=gpurun_out/r2z_bench_lines.jsonl
: > $OUT
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2z_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_gputests.log
python bench.py --steps 5 --warmup 3 2> gpurun_out/r2z_c3.err | tail -1 >> $OUT
python bench.py --config c5 --steps 3 --warmup 3 2> gpurun_out/r2z_c5.err | tail -1 >> $OUT
python bench.py --config c4 --steps 3 --warmup 1 2> gpurun_out/r2z_c4.err | tail -1 >> $OUT
python scripts/time_firconv.py 2> /dev/null | tail -1 >> $OUT
python scripts/time_tdsolver.py 2> /dev/null | tail -1 >> $OUT
CNLD_B200_TD_BLOCK=1 python scripts/time_tdsolver.py --no-cpu 2> /dev/null | tail -1 >> $OUT
CNLD_B200_TD_CLUSTER=0 python scripts/time_tdsolver.py --no-cpu 2> /dev/null | tail -1 >> $OUT
python scripts/time_hmv.py 16 19 33.4 1,96 2> /dev/null | grep ywarps >> $OUT
wc -l $OUT
# ncu (each command has just run without ncu above)
# blocks of 16: launches matching the regex = 1 (initial state) + 4050 steps + 254 history passes before the timed steps
ncu --set full --clock-control none --import-source on -k regex:"k_fir_far_mma|k_td_step_cl" --launch-skip 4319 --launch-count 3 \
    -o gpurun_out/r2z_prof_td -f python scripts/time_tdsolver.py --steps 60 --no-cpu > gpurun_out/r2z_ncu1.log 2>&1
CNLD_B200_TD_BLOCK=1 ncu --set full --clock-control none --import-source on -k regex:"k_fir_conv" --launch-skip 4070 --launch-count 1 \
    -o gpurun_out/r2z_prof_conv -f python scripts/time_tdsolver.py --steps 60 --no-cpu > gpurun_out/r2z_ncu2.log 2>&1
python scripts/time_hmv1.py > gpurun_out/r2z_hmv1_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_hmv1" --launch-skip 4 --launch-count 2 \
    -o gpurun_out/r2z_prof_hmv1 -f python scripts/time_hmv1.py > gpurun_out/r2z_ncu3.log 2>&1
ls -la gpurun_out | grep r2z_
