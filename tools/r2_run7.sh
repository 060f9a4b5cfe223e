set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest7.log; tail -8 gpurun_out/r2_pytest7.log
rm -f gpurun_out/r2_modes7.jsonl
for p in f16c8 bf16x3 f16; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision $p >> gpurun_out/r2_modes7.jsonl 2>> gpurun_out/r2_modes7.err
done
VETO_ATTENTION_SPLIT=0 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision f16c8 >> gpurun_out/r2_modes7.jsonl 2>> gpurun_out/r2_modes7.err
cat gpurun_out/r2_modes7.jsonl; tail -3 gpurun_out/r2_modes7.err
