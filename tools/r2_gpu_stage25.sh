#!/bin/bash
# single-thread MMA issue (forward, backward) + several TMA stores in flight: parity, knobs, timing
mkdir -p gpurun_out
L=gpurun_out/stage25.log
: > $L
run() { echo "== $*" >> $L; timeout 400 "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
if CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 64 512 97 97 bf16; then
CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 64 512 97 97 fp32
CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so CCA_B200_BF16_NATIVE=1 run python tools/r2_probe.py parity 1 32 128 113 200 bf16
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu
run python tools/r2_probe.py time 8 64 512 97 97 bf16
run python tools/r2_probe.py time 8 64 512 97 97 fp32
run python tools/r2_probe.py time 8 64 512 65 65 bf16
fi
grep -E "^\{\"mode|rc=[^0]|passed|failed|rror" $L | cut -c1-420
