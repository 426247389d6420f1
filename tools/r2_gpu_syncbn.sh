#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/r2_gpu_syncbn.py > gpurun_out/syncbn.json 2> gpurun_out/syncbn.err; echo "rc=$?" >> gpurun_out/syncbn.err
grep "^{" gpurun_out/syncbn.json | cut -c1-900; tail -3 gpurun_out/syncbn.err | cut -c1-300
