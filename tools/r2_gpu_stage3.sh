#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/stage3.log
: > $L
run() { echo "== $*" >> $L; timeout 240 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for dt in fp32 bf16; do
  run python tools/r2_probe.py parity 2 16 64 5 6 $dt
  run python tools/r2_probe.py parity 8 64 512 97 97 $dt
  run python tools/r2_probe.py parity 1 64 512 129 129 $dt
  run python tools/r2_probe.py parity 1 32 128 113 200 $dt
done
run env CCA_B200_LAG=0 python tools/r2_probe.py parity 8 64 512 97 97 fp32
run env CCA_B200_LAG=1 CCA_B200_DELTA=0 python tools/r2_probe.py parity 8 64 512 97 97 fp32
for dt in fp32 bf16; do
  run python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_L2HINT=0 python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_LAG=0 python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_LAG=1 python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_DELTA=0 python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_LAG=1 CCA_B200_L2HINT=0 python tools/r2_probe.py time 8 64 512 97 97 $dt
done
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python tools/r2_probe.py time 8 64 512 129 129 fp32
run python tools/r2_probe.py time 8 64 512 193 193 fp32
ncu --set full --clock-control none --import-source on -k regex:cca_ -s 4 -c 4 -o gpurun_out/r02b_op python tools/run_op.py 3 > gpurun_out/ncu_r02b.log 2>&1
grep -E "^\{|rc=|==" $L | cut -c1-420
