#!/bin/bash
# A/B of two builds on ONE box after the kernel tests:  tools/gpu_ab_lib.sh <pytest -k selection> <libA.so> <libB.so>
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "$1" 2>&1 | tail -3
bash tools/ab.sh $2 $3
