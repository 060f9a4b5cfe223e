set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes14.*
for r in 0 1 0 1; do
VETO_RESIDUAL_OPERAND=$r timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes14.jsonl 2>> gpurun_out/r2_modes14.err
done
VETO_RESIDUAL_OPERAND=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "f16c8 or forward or logits" 2>&1 | tail -5 > gpurun_out/r2_pytest14.log
cat gpurun_out/r2_modes14.jsonl; tail -3 gpurun_out/r2_modes14.err; cat gpurun_out/r2_pytest14.log
