#!/bin/bash
# Final round-2 numbers on one B200 (run under gpurun): the three bench configs + auxiliary timings -> gpurun_out/r2_bench_lines.jsonl
set -x
OUT=gpurun_out/r2_bench_lines.jsonl
: > $OUT
python bench.py 2> gpurun_out/r2_c3.err | tail -1 >> $OUT
python bench.py --general --no-cpu-baseline 2> /dev/null | tail -1 >> $OUT
python bench.py --config c5 --steps 3 --warmup 3 2> gpurun_out/r2_c5.err | tail -1 >> $OUT
python bench.py --config c4 --steps 3 --warmup 1 2> gpurun_out/r2_c4.err | tail -1 >> $OUT
python bench.py --impl reference --steps 1 --warmup 0 2> /dev/null | tail -1 >> $OUT
python scripts/time_generate_db.py 12 1 2> /dev/null | tail -1 >> $OUT
python scripts/time_generate_db.py 24 0 2> /dev/null | tail -1 >> $OUT
python scripts/time_firconv.py 2> /dev/null | tail -1 >> $OUT
python scripts/time_pfield.py 303104 3 2> /dev/null | tail -1 >> $OUT
python scripts/time_hmv.py 16 19 8.6 1,96 2> /dev/null | grep ywarps >> $OUT
python scripts/time_hsweep.py 8.6 0,4 2> /dev/null | grep coarse >> $OUT
wc -l $OUT
