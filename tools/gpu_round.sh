#!/bin/bash
# One gpurun call: parity tests, bench (+ per-kernel table), ncu launch list, one full ncu capture of the conv kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 5 --dump gpurun_out/per_kernel.json > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
if [ "$1" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --min-warm-sec 0 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_shift|conv_igemm|maxpool|stem' -s 150 -c 26 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --min-warm-sec 0 > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
