set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes30.*
for v in 0 1 0 1; do
VETO_QKV_ITEM_LAYOUT=$v timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes30.jsonl 2>> gpurun_out/r2_modes30.err
done
cat gpurun_out/r2_modes30.jsonl; tail -5 gpurun_out/r2_modes30.err
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
