#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/stage16.log
: > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for h in 1 2 0; do
  for lag in 1 0; do
    run env CCA_B200_L2HINT=$h CCA_B200_LAG=$lag python tools/r2_probe.py time 8 64 512 97 97 fp32
  done
done
run env CCA_B200_L2HINT=2 python tools/r2_probe.py time 8 64 512 97 97 bf16
run env CCA_B200_L2HINT=2 python tools/r2_probe.py parity 8 64 512 97 97 fp32
grep -E "^\{\"mode|rc=[^0]" $L | cut -c1-330
