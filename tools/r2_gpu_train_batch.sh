#!/bin/bash
# synthetic ResNet101+RCCA train step on one GPU at batch 8 / 4 / 2: is the step time proportional to the batch?
mkdir -p gpurun_out
: > gpurun_out/train_batch.jsonl
for b in 8 4 2; do timeout 300 python -m harness.train_synth --batch $b --tf32 --steps 4 --warmup 3 >> gpurun_out/train_batch.jsonl 2>> gpurun_out/train_batch.err; done
cut -c1-330 gpurun_out/train_batch.jsonl
