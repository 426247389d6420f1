#!/bin/bash
# End-of-round evidence run (release build in ccnet_b200/lib): tests, smoke, bench lines, ncu captures (fp32 + bf16), launch list,
# resolution sweep, parity report, module profile, output-path microbenchmarks, compute-sanitizer.  Everything lands in
# gpurun_out/final/ and is copied to profiles/r02_* afterwards.
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
cat ccnet_b200/lib/flavour.txt >> $O/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_line.json 2> $O/bench_err.txt; echo "bench rc=$?" >> $O/bench_err.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_line_reference_arm.json 2>> $O/bench_err.txt
ncu --set full --clock-control none --import-source on -k regex:cca_ -s 4 -c 4 -f -o $O/r02_op python tools/run_op.py 3 > $O/ncu_op.log 2>&1
ncu --set full --clock-control none -k regex:cca_ -s 4 -c 4 -f -o $O/r02_op_bf16 python tools/run_op.py 3 bf16 > $O/ncu_op_bf16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-train > /dev/null 2>&1
timeout 600 python tools/sweep.py > $O/resolution_sweep.jsonl 2>> $O/bench_err.txt
timeout 300 python tools/module_profile.py > $O/module_profile.txt 2>&1
( timeout 120 tools/tma_red_bench > $O/tma_red_bench.jsonl 2>&1; timeout 120 tools/tma_mix_bench > $O/tma_mix_bench.jsonl 2>&1 )
for dt in fp32 bf16; do
  for shp in "2 16 64 5 6" "8 64 512 97 97" "2 64 512 65 65" "1 64 512 129 129" "1 32 128 113 200" "1 64 512 193 193"; do
    timeout 400 python tools/r2_probe.py parity $shp $dt 2 >> $O/parity_report.jsonl 2>> $O/bench_err.txt
  done
done
CCA_B200_BF16_NATIVE=1 timeout 400 python tools/r2_probe.py parity 1 32 128 113 200 bf16 2 >> $O/parity_report.jsonl 2>> $O/bench_err.txt
for tool in memcheck racecheck; do
  for dt in fp32 bf16; do
    echo "== compute-sanitizer --tool $tool: parity 2 16 64 20 33 $dt" >> $O/sanitizer_$tool.txt
    timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/r2_probe.py parity 2 16 64 20 33 $dt >> $O/sanitizer_$tool.txt 2>&1
    echo "rc=$?" >> $O/sanitizer_$tool.txt
  done
done
tail -3 $O/pytest_gpu.txt; cat $O/smoke.txt | tail -2; cut -c1-600 $O/bench_line.json; tail -5 $O/bench_err.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=" $O/sanitizer_*.txt | cut -c1-200
