set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes27.*
for v in 0 1 0 1; do
if [ $v = 1 ]; then export VETO_FF1_BN256=1; else unset VETO_FF1_BN256; fi
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes27.jsonl 2>> gpurun_out/r2_modes27.err
done
cat gpurun_out/r2_modes27.jsonl; tail -5 gpurun_out/r2_modes27.err
VETO_FF1_BN256=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "relation_logits or gemm_tcgen05 or meet" 2>&1 | tail -3
