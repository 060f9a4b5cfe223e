set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes23.*
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes23.jsonl 2>> gpurun_out/r2_modes23.err
cp veto_b200/lib/libveto_b200.so /tmp/lib_default.so
VETO_NVCC_DEFINES="-DVETO_TC2_RES_STAGES=2" timeout 600 python -c "
from veto_b200 import build; build.build_library_locked(force=True)"
for i in 1 2; do
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes23.jsonl 2>> gpurun_out/r2_modes23.err
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "relation_logits or gemm_tcgen05" 2>&1 | tail -3
cp /tmp/lib_default.so veto_b200/lib/libveto_b200.so
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes23.jsonl 2>> gpurun_out/r2_modes23.err
cat gpurun_out/r2_modes23.jsonl; tail -5 gpurun_out/r2_modes23.err
