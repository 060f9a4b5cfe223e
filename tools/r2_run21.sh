set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes21.*
for d in 0 2 3 4; do
VETO_GEMM_DIAG=$d timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes21.jsonl 2>> gpurun_out/r2_modes21.err
done
cat gpurun_out/r2_modes21.jsonl; tail -5 gpurun_out/r2_modes21.err
