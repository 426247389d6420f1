#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/stage14_bench.json 2> gpurun_out/stage14_err.txt
echo "rc=$?" >> gpurun_out/stage14_err.txt
tail -5 gpurun_out/stage14_err.txt
cut -c1-3000 gpurun_out/stage14_bench.json
