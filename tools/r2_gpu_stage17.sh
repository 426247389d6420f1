#!/bin/bash
# backward with dQ/dK accumulators of their own: parity, timing fp32 / bf16 / tiled
mkdir -p gpurun_out
L=gpurun_out/stage17.log
: > $L
run() { echo "== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "backward or knobs or grad"
run python tools/r2_probe.py time 8 64 512 97 97 fp32
run python tools/r2_probe.py time 8 64 512 97 97 bf16
run python tools/r2_probe.py time 8 64 512 65 65 fp32
run python tools/r2_probe.py time 2 64 512 129 129 fp32
run python tools/r2_probe.py time 1 64 512 193 193 fp32
run python tools/r2_probe.py parity 8 64 512 97 97 fp32
grep -E "^\{\"mode|rc=[^0]|passed|failed" $L | cut -c1-330
