#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/stage8.log
: > $L
run() { echo "== $*" >> $L; timeout 400 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for dt in bf16 fp32; do
  run python tools/r2_probe.py parity 2 16 64 5 6 $dt
  run python tools/r2_probe.py parity 8 64 512 97 97 $dt
  run python tools/r2_probe.py parity 1 64 512 129 129 $dt
  run python tools/r2_probe.py parity 1 32 128 113 200 $dt
  run python tools/r2_probe.py parity 1 64 512 193 193 $dt 1
done
run python -m pytest tests/test_gpu_parity.py -q -x -k "launch_knobs or full_batch or fused_module"
for dt in fp32 bf16; do
  run python tools/r2_probe.py time 8 64 512 97 97 $dt
done
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python tools/r2_probe.py time 8 64 512 129 129 fp32
run python tools/r2_probe.py time 8 64 512 193 193 fp32
run python tools/r2_probe.py time 8 64 512 129 129 bf16
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cca_ -s 8 -c 8 --csv --log-file gpurun_out/r02c_launches_bf16.csv python tools/r2_timeline.py bf16 > /dev/null 2>&1
grep -E "^\{\"mode|rc=|==|passed|failed" $L | cut -c1-400
grep -E "cca_tc" gpurun_out/r02c_launches_bf16.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
