#!/bin/bash
mkdir -p gpurun_out
CCA_B200_LIB=$PWD/ccnet_b200/lib_tl/libcca_b200.so timeout 300 python tools/r2_timeline.py bf16 > gpurun_out/stage15.log 2>&1
bash tools/sanitize.sh >> gpurun_out/stage15.log 2>&1
tail -30 gpurun_out/stage15.log | cut -c1-250
