set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes24.*
for v in 0 1 0 1; do
VETO_FF1_SHALLOW=$v timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes24.jsonl 2>> gpurun_out/r2_modes24.err
done
cat gpurun_out/r2_modes24.jsonl; tail -5 gpurun_out/r2_modes24.err
VETO_FF1_SHALLOW=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "relation_logits or gemm_tcgen05" 2>&1 | tail -3
