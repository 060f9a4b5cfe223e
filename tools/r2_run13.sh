set -x
mkdir -p gpurun_out
for p in 0 64 56 48 37; do
VETO_GEMM_PAIRS=$p timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes13.jsonl 2>> gpurun_out/r2_modes13.err
done
VETO_GEMM_PAIRS=56 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision bf16x3 >> gpurun_out/r2_modes13.jsonl 2>> gpurun_out/r2_modes13.err
VETO_GEMM_PAIRS=56 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16 >> gpurun_out/r2_modes13.jsonl 2>> gpurun_out/r2_modes13.err
cat gpurun_out/r2_modes13.jsonl; tail -3 gpurun_out/r2_modes13.err
