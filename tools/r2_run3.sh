# round 2, GPU call 3: GPU suite, probes per precision mode, ncu of the f16c8 / f16 GEMMs and the new gather kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_pytest3.log; tail -12 gpurun_out/r2_pytest3.log
rm -f gpurun_out/r2_modes.jsonl
for p in bf16x3 f16c8 f16; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision $p >> gpurun_out/r2_modes.jsonl 2>> gpurun_out/r2_modes.err
done
cat gpurun_out/r2_modes.jsonl; tail -3 gpurun_out/r2_modes.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_kernel' -c 6 \
  -o gpurun_out/r2_gemm_f16c8 -f python tools/infer_probe.py --images 4 --once --precision f16c8 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_kernel' -c 6 \
  -o gpurun_out/r2_gemm_f16 -f python tools/infer_probe.py --images 4 --once --precision f16 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_gather|attention_mma2' -c 4 \
  -o gpurun_out/r2_gather_v2 -f python tools/infer_probe.py --images 32 --once --precision f16c8 > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log
