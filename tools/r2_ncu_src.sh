set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2' -c 2 \
  -o gpurun_out/r2_qkv_out_src -f python tools/infer_probe.py --images 4 --once --precision f16c8 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/r2_qkv_out_src.ncu-rep; du -sh gpurun_out
