#!/bin/bash
# Mid-round evidence run (release build): tests, smoke, bench lines, ncu captures, GPU-only sweep.
mkdir -p gpurun_out/eva
O=gpurun_out/eva
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
cat ccnet_b200/lib/flavour.txt >> $O/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_line.json 2> $O/bench_err.txt; echo "bench rc=$?" >> $O/bench_err.txt
ncu --set full --clock-control none --import-source on -k regex:cca_ -s 4 -c 4 -o $O/r02_op python tools/run_op.py 3 > $O/ncu_op.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-train > /dev/null 2>&1
timeout 600 python tools/sweep.py --no-cpu > $O/resolution_sweep.jsonl 2>> $O/bench_err.txt
tail -3 $O/pytest_gpu.txt; cat $O/smoke.txt | tail -2; cut -c1-1500 $O/bench_line.json; tail -5 $O/bench_err.txt; cat $O/resolution_sweep.jsonl
