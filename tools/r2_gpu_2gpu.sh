#!/bin/bash
# two-GPU check of the bench contract (the driver's scaling run launches it the same way): own arm and reference arm
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
cut -c1-400 gpurun_out/bench_2gpu.json; tail -4 gpurun_out/bench_2gpu.err; cut -c1-200 gpurun_out/bench_2gpu_ref.json
