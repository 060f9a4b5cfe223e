set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes25.* gpurun_out/r2_pytest25.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_pytest25.log
cat gpurun_out/r2_pytest25.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
for d in 0 0; do
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes25.jsonl 2>> gpurun_out/r2_modes25.err
done
cat gpurun_out/r2_modes25.jsonl; tail -5 gpurun_out/r2_modes25.err
