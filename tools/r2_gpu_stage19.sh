#!/bin/bash
# ncu --set full of the channel-major forward values kernel (one launch), with source-level sampling
mkdir -p gpurun_out
export CCA_B200_FWDT=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fwdt -s 2 -c 1 -f -o gpurun_out/r02_fwdt python tools/run_op.py 3 > gpurun_out/stage19.log 2>&1
echo "rc=$?" >> gpurun_out/stage19.log
tail -5 gpurun_out/stage19.log
