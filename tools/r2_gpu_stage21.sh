#!/bin/bash
# statistics pre-pass: single-pass online softmax with tensor-memory prefetch -- forward parity tests, timing, launch list
mkdir -p gpurun_out
L=gpurun_out/stage21.log
: > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
CCA_B200_LIB=ccnet_b200/lib_dbg/libcca_b200.so run python tools/r2_probe.py parity 2 64 512 97 97 fp32
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward or golden or peaky or bf16 or module"
run python tools/r2_probe.py time 8 64 512 97 97 fp32
run python tools/r2_probe.py time 8 64 512 97 97 bf16
run python tools/r2_probe.py time 8 64 512 65 65 fp32
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cca_tc -c 12 --csv --log-file gpurun_out/stage21_launches.csv python tools/run_op.py 3 >> $L 2>&1
grep -E "^\{\"mode|rc=[^0]|passed|failed|rror" $L | cut -c1-420; grep -E "stats|fwd" gpurun_out/stage21_launches.csv | tail -6 | cut -c1-300
