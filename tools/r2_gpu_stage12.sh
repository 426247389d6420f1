#!/bin/bash
mkdir -p gpurun_out
CCA_B200_LIB=$PWD/ccnet_b200/lib_tl/libcca_b200.so timeout 300 python tools/r2_timeline.py fp32 > gpurun_out/stage12.log 2>&1
tail -2 gpurun_out/stage12.log
