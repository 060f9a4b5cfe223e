# round 2, GPU call 5: LayerNorm fusion + compile-time epilogues: GPU suite, probes with / without the fusion
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2_pytest5.log; tail -12 gpurun_out/r2_pytest5.log
rm -f gpurun_out/r2_modes5.jsonl
for p in f16c8 bf16x3 f16; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision $p >> gpurun_out/r2_modes5.jsonl 2>> gpurun_out/r2_modes5.err
done
VETO_LN_FUSION=0 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision f16c8 >> gpurun_out/r2_modes5.jsonl 2>> gpurun_out/r2_modes5.err
VETO_LN_FUSION=0 VETO_GEMM_GENERIC_EPI=1 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision f16c8 >> gpurun_out/r2_modes5.jsonl 2>> gpurun_out/r2_modes5.err
cat gpurun_out/r2_modes5.jsonl; tail -3 gpurun_out/r2_modes5.err
