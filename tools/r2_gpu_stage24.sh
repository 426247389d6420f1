#!/bin/bash
# in-kernel timelines (CTA 0) of the statistics, values and backward kernels, fp32 and bf16
mkdir -p gpurun_out
export CCA_B200_LIB=ccnet_b200/lib_tl/libcca_b200.so
timeout 200 python tools/r2_timeline.py fp32 > gpurun_out/stage24.log 2>&1
timeout 200 python tools/r2_timeline.py bf16 >> gpurun_out/stage24.log 2>&1
tail -4 gpurun_out/stage24.log
