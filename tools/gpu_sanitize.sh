#!/bin/bash
# compute-sanitizer over the conv-kernel parity tests (mbarrier rings, TMEM hand-offs, stream-K flags): racecheck, synccheck, memcheck
mkdir -p gpurun_out
SEL='conv_fp16_kernels_vs_oracle or stream_k or stride2 or stacked_conv or test_op_vs_reference_golden or readme or pointwise or small_first_layer or fp32_on_the_tensor_pipe or register_transpose or avgpool_3x3'
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool ===" 
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "exit code $?"
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" gpurun_out/r2_sanitizer_$tool.log | sort | uniq -c | sort -rn | head -12
done
