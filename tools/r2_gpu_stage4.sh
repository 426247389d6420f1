#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/stage4.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "qkv or fused_module or torch_ops" >> $L 2>&1
echo "rc=$?" >> $L
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "launch_knobs or full_batch or golden" >> $L 2>&1
echo "rc=$?" >> $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-train >> $L 2>&1
echo "rc=$?" >> $L
tail -60 $L | cut -c1-3000
