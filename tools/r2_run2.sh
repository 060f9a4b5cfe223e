# round 2, GPU call 2: the whole GPU suite (new: f16c8 / f16 modes, rebuilt gather + backward, MEET sampling kernel,
# full-size fixtures, oversized post-processing), then the inference probe per precision mode
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25
for p in bf16x3 f16c8 f16; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 7976 --steps 2 --warmup 1 --precision $p >> gpurun_out/r2_modes.jsonl 2>> gpurun_out/r2_modes.err
done
cat gpurun_out/r2_modes.jsonl; tail -3 gpurun_out/r2_modes.err
