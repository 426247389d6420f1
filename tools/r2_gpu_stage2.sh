#!/bin/bash
mkdir -p gpurun_out
./tools/tma_red_bench > gpurun_out/tma_red_bench.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:cca_ -s 4 -c 4 -o gpurun_out/r02a_op python tools/run_op.py 3 > gpurun_out/ncu_r02a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cca_ -s 8 -c 8 --csv --log-file gpurun_out/r02a_launches.csv python tools/run_op.py 4 > /dev/null 2>&1
cat gpurun_out/tma_red_bench.jsonl
tail -3 gpurun_out/ncu_r02a.log
cat gpurun_out/r02a_launches.csv | tail -12
