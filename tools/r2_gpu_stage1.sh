#!/bin/bash
# first contact of the round-2 kernels with the GPU: small -> large, one process per case, everything logged
mkdir -p gpurun_out
L=gpurun_out/stage1.log
: > $L
run() { echo "== $*" >> $L; timeout 180 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $L 2>&1
for dt in fp32 bf16; do
  run python tools/r2_probe.py parity 2 16 64 5 6 $dt
  run python tools/r2_probe.py parity 1 64 512 97 97 $dt
  run python tools/r2_probe.py parity 2 32 256 20 97 $dt
  run python tools/r2_probe.py parity 8 64 512 97 97 $dt
  run python tools/r2_probe.py parity 1 64 512 129 129 $dt
  run python tools/r2_probe.py parity 1 32 128 113 200 $dt
  run python tools/r2_probe.py parity 1 64 512 193 193 $dt 1
done
run env CCA_B200_DELTA=0 python tools/r2_probe.py parity 8 64 512 97 97 fp32
run env CCA_B200_ZERO_AHEAD=2 python tools/r2_probe.py parity 8 64 512 97 97 fp32
for dt in fp32 bf16; do
  run python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_DELTA=0 python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_ZERO_AHEAD=2 python tools/r2_probe.py time 8 64 512 97 97 $dt
  run env CCA_B200_PDL=0 python tools/r2_probe.py time 8 64 512 97 97 $dt
done
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python tools/r2_probe.py time 8 64 512 129 129 fp32
run python tools/r2_probe.py time 8 64 512 193 193 fp32
grep -E "^\{|rc=|==" $L | tail -120
