#!/bin/bash
# channel-major forward values kernel (cca_tc_fwdt.cu): parity on the debug build (bounded spins), then timing A/B on the release build
mkdir -p gpurun_out
L=gpurun_out/stage18.log
: > $L
run() { echo "== $*" >> $L; timeout 90 "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
export CCA_B200_FWDT=1
export CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so
if run python tools/r2_probe.py parity 2 64 512 97 97 fp32; then
run python tools/r2_probe.py parity 8 64 512 97 97 fp32
run python tools/r2_probe.py parity 1 64 512 129 129 fp32
run python tools/r2_probe.py parity 2 16 256 40 150 fp32
unset CCA_B200_LIB
run python tools/r2_probe.py time 8 64 512 97 97 fp32
CCA_B200_FWDT=0 run python tools/r2_probe.py time 8 64 512 97 97 fp32
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python tools/r2_probe.py time 2 64 512 129 129 fp32
run python tools/r2_probe.py time 1 64 512 193 193 fp32
fi
grep -E "^\{\"mode|rc=[^0]|passed|failed|rror" $L | cut -c1-420
