# GPU round for the depth backbone: parity tests, timing, ncu launch list (pass "ncu" as $1)
set -x
timeout 600 python -m pytest tests/test_depth_backbone.py -q -m gpu 2>&1 | tail -5
timeout 300 python tools/depth_bench.py 2>&1 | tail -1 | tee gpurun_out/depth_bench.json
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/depth_launches.csv python tools/depth_bench.py --profile > gpurun_out/depth_ncu.log 2>&1
fi
