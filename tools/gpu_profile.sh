#!/bin/bash
# Round profile: ncu launch list of a short bench run + one full capture of every kernel of one steady-state step.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --min-warm-sec 0 > gpurun_out/ncu_bench.log 2>&1
# the executor build launches ~120 one-off packing kernels; the last eager forward before capture starts after them
ncu --set full --clock-control none --import-source on -k regex:'conv_shift|conv_stack|conv_igemm|stem_pool|gap_dense|pooled_dense' -s 36 -c 18 -f \
    -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline --min-warm-sec 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out | head -20
