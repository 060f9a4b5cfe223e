set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes15.* gpurun_out/r2_pytest15*.log
VETO_GEMM_CLUSTER4=2 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_pytest15a.log
cat gpurun_out/r2_pytest15a.log
for r in 0 1 0 1; do
VETO_GEMM_CLUSTER4=$r timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes15.jsonl 2>> gpurun_out/r2_modes15.err
done
VETO_GEMM_CLUSTER4=1 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision bf16x3 >> gpurun_out/r2_modes15.jsonl 2>> gpurun_out/r2_modes15.err
cat gpurun_out/r2_modes15.jsonl; tail -5 gpurun_out/r2_modes15.err
