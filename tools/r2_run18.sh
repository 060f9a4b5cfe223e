set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes18.* gpurun_out/r2_pytest18*.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_pytest18.log
cat gpurun_out/r2_pytest18.log
for r in 1 2; do
VETO_LN_STATS_KERNEL=1 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes18.jsonl 2>> gpurun_out/r2_modes18.err
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes18.jsonl 2>> gpurun_out/r2_modes18.err
done
cat gpurun_out/r2_modes18.jsonl; tail -5 gpurun_out/r2_modes18.err
