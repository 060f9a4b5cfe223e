set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes26.*
cp veto_b200/lib/libveto_b200.so /tmp/lib_new.so
for v in prev new prev new; do
if [ $v = prev ]; then cp veto_b200/lib/libveto_b200_prev.so veto_b200/lib/libveto_b200.so; else cp /tmp/lib_new.so veto_b200/lib/libveto_b200.so; fi
touch veto_b200/lib/libveto_b200.so
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes26.jsonl 2>> gpurun_out/r2_modes26.err
done
cat gpurun_out/r2_modes26.jsonl; tail -5 gpurun_out/r2_modes26.err
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
