#!/bin/bash
# bf16 forward: second epilogue group (four of the idle converter warps) -- parity, knobs, timing
mkdir -p gpurun_out
L=gpurun_out/stage23.log
: > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
if CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 64 512 97 97 bf16; then
CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 16 192 33 47 bf16
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bf16 or knobs or noise"
run python tools/r2_probe.py time 8 64 512 97 97 bf16
run python tools/r2_probe.py time 8 64 512 65 65 bf16
run python tools/r2_probe.py time 8 64 512 97 97 fp32
fi
grep -E "^\{\"mode|rc=[^0]|passed|failed|rror" $L | cut -c1-420
