set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest10.log; tail -8 gpurun_out/r2_pytest10.log
rm -f gpurun_out/r2_modes10.jsonl
for p in f16c8 bf16x3; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision $p >> gpurun_out/r2_modes10.jsonl 2>> gpurun_out/r2_modes10.err
done
VETO_RESIDUAL_OPERAND=0 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes10.jsonl 2>> gpurun_out/r2_modes10.err
cat gpurun_out/r2_modes10.jsonl; tail -3 gpurun_out/r2_modes10.err
