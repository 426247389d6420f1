#!/bin/bash
# statistics pre-pass with four statistics groups: forward parity tests, timing, launch list
mkdir -p gpurun_out
L=gpurun_out/stage22.log
: > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
if CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 64 512 97 97 fp32; then
CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 64 512 97 97 bf16
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward or golden or peaky or bf16 or module or noise or channel_major"
run python tools/r2_probe.py time 8 64 512 97 97 fp32
run python tools/r2_probe.py time 8 64 512 97 97 bf16
run python tools/r2_probe.py time 8 64 512 65 65 fp32
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cca_tc -c 6 --csv --log-file gpurun_out/stage22_launches.csv python tools/run_op.py 2 >> $L 2>&1
fi
grep -E "^\{\"mode|rc=[^0]|passed|failed|rror" $L | cut -c1-420; grep -E "stats" gpurun_out/stage22_launches.csv | tail -2 | cut -c150-330
