#!/bin/bash
# full ncu capture of single conv launches through tools/role_profile2.py:  [NCU_EXTRA="..."] tools/ncu_one.sh <keys> <out-name>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on $NCU_EXTRA -k regex:'conv_shift|conv_igemm|conv_stack' -s 5 -c 1 -o gpurun_out/$2 -f \
    python tools/role_profile2.py $1 > gpurun_out/$2.log 2>&1
tail -3 gpurun_out/$2.log
