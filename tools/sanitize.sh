#!/bin/bash
# compute-sanitizer runs of the tensor-core forward + backward at a small covered shape: memcheck, racecheck,
# initcheck, synccheck.  Output -> gpurun_out/sanitizer_<tool>.txt (summaries are committed under profiles/).
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  for dt in fp32 bf16; do
    echo "== compute-sanitizer --tool $tool: parity 2 16 64 20 33 $dt" >> gpurun_out/sanitizer_$tool.txt
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/r2_probe.py parity 2 16 64 20 33 $dt >> gpurun_out/sanitizer_$tool.txt 2>&1
    echo "rc=$?" >> gpurun_out/sanitizer_$tool.txt
  done
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|max_abs_err" gpurun_out/sanitizer_$tool.txt | cut -c1-300
done
