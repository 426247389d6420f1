#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/stage13.log
: > $L
run() { echo "== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for dt in fp32 bf16; do
  run python tools/r2_probe.py parity 2 16 64 5 6 $dt
  run python tools/r2_probe.py parity 8 64 512 97 97 $dt
  run python tools/r2_probe.py parity 1 64 512 129 129 $dt
done
for dt in fp32 bf16; do
  run python tools/r2_probe.py time 8 64 512 97 97 $dt
done
run python tools/r2_probe.py time 8 64 512 129 129 fp32
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python -m pytest tests/test_gpu_parity.py -q
CCA_B200_LIB=$PWD/ccnet_b200/lib_tl/libcca_b200.so timeout 300 python tools/r2_timeline.py fp32 >> $L 2>&1
grep -E "^\{\"mode|rc=|==|passed|failed|Error" $L | cut -c1-330
