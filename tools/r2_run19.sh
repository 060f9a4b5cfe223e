set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes19.*
for d in 0 1; do
VETO_GEMM_DIAG=$d timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes19.jsonl 2>> gpurun_out/r2_modes19.err
done
VETO_GEMM_DIAG=1 timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision bf16x3 >> gpurun_out/r2_modes19.jsonl 2>> gpurun_out/r2_modes19.err
cat gpurun_out/r2_modes19.jsonl; tail -5 gpurun_out/r2_modes19.err
