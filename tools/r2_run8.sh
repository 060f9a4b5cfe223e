set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest8.log; tail -5 gpurun_out/r2_pytest8.log
rm -f gpurun_out/r2_modes8.jsonl
for p in f16c8 bf16x3; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision $p >> gpurun_out/r2_modes8.jsonl 2>> gpurun_out/r2_modes8.err
done
cat gpurun_out/r2_modes8.jsonl; tail -3 gpurun_out/r2_modes8.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_split|gemm_tc2' -c 12 \
  -o gpurun_out/r2_layer_f16c8 -f python tools/infer_probe.py --images 4 --once --precision f16c8 > gpurun_out/ncu8.log 2>&1
tail -2 gpurun_out/ncu8.log
