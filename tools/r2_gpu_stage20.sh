#!/bin/bash
# channel-major forward values kernel, lean epilogue: parity (debug build), timing, ncu capture
mkdir -p gpurun_out
L=gpurun_out/stage20.log
: > $L
run() { echo "== $*" >> $L; timeout 90 "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
export CCA_B200_FWDT=1
export CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so
if run python tools/r2_probe.py parity 8 64 512 97 97 fp32; then
run python tools/r2_probe.py parity 1 64 512 129 129 fp32
unset CCA_B200_LIB
run python tools/r2_probe.py time 8 64 512 97 97 fp32
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python tools/r2_probe.py time 2 64 512 129 129 fp32
run python tools/r2_probe.py time 1 64 512 193 193 fp32
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fwdt -s 2 -c 1 -f -o gpurun_out/r02_fwdt python tools/run_op.py 3 >> $L 2>&1
fi
grep -E "^\{\"mode|rc=[^0]|passed|failed|rror" $L | cut -c1-420
