set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes28.* gpurun_out/r2_pytest28.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_pytest28.log
cat gpurun_out/r2_pytest28.log
for d in 0 0 0; do
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes28.jsonl 2>> gpurun_out/r2_modes28.err
done
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision bf16x3 >> gpurun_out/r2_modes28.jsonl 2>> gpurun_out/r2_modes28.err
cat gpurun_out/r2_modes28.jsonl; tail -5 gpurun_out/r2_modes28.err
