#!/bin/bash
mkdir -p gpurun_out
CCA_B200_LIB=$PWD/ccnet_b200/lib_tl/libcca_b200.so timeout 300 python tools/r2_timeline.py bf16 > gpurun_out/stage7.log 2>&1
timeout 300 python tools/module_profile.py > gpurun_out/module_profile.txt 2>&1
head -30 gpurun_out/module_profile.txt | cut -c1-160
