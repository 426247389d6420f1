#!/bin/bash
mkdir -p gpurun_out
export CCA_B200_LIB=$PWD/ccnet_b200/lib_tl/libcca_b200.so
./tools/tma_mix_bench > gpurun_out/tma_mix_bench.jsonl 2>&1
timeout 300 python tools/r2_timeline.py fp32 > gpurun_out/stage5.log 2>&1
timeout 300 python tools/r2_timeline.py bf16 >> gpurun_out/stage5.log 2>&1
cat gpurun_out/stage5.log | tail -5
cat gpurun_out/tma_mix_bench.jsonl
timeout 300 python tools/module_profile.py > gpurun_out/module_profile.txt 2>&1; tail -45 gpurun_out/module_profile.txt
