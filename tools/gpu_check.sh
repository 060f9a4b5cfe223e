#!/usr/bin/env bash
# First-contact / regression diagnostics on the GPU box: every stage in its own process with a timeout,
# logs under gpurun_out/.  Usage: tools/gpu_check.sh [stage ...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES=("$@")
[ ${#STAGES[@]} -eq 0 ] && STAGES=(simt rows pairs gather tc head:fp32 head:bf16x3 head:bf16 speed:bf16 speed:bf16x3 speed:fp32)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_gpu.txt 2>&1
: > gpurun_out/diag.jsonl
for s in "${STAGES[@]}"; do
  echo "=== $s" | tee -a gpurun_out/diag.log
  timeout 300 python tests/gpu_diag.py "$s" > gpurun_out/diag_stage.out 2> gpurun_out/diag_stage.err
  rc=$?
  grep '^{' gpurun_out/diag_stage.out >> gpurun_out/diag.jsonl
  cat gpurun_out/diag_stage.out | tee -a gpurun_out/diag.log
  echo "rc=$rc" | tee -a gpurun_out/diag.log
  if [ $rc -ne 0 ]; then tail -n 30 gpurun_out/diag_stage.err | tee -a gpurun_out/diag.log; fi
done
