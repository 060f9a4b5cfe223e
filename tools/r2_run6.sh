set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest6.log; tail -8 gpurun_out/r2_pytest6.log
rm -f gpurun_out/r2_modes6.jsonl
for p in f16c8 bf16x3; do
  timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 2 --warmup 1 --precision $p >> gpurun_out/r2_modes6.jsonl 2>> gpurun_out/r2_modes6.err
done
cat gpurun_out/r2_modes6.jsonl; tail -3 gpurun_out/r2_modes6.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_gather' -c 2 \
  -o gpurun_out/r2_gather_v3 -f python tools/infer_probe.py --images 32 --once --precision f16c8 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log
