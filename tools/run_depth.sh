set -x
timeout 600 python -m pytest tests/test_depth_backbone.py -q -m gpu 2>&1 | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 300 python tools/depth_bench.py 2>&1 | tail -2 | tee gpurun_out/depth_bench.json
timeout 600 python bench.py --no-inference > gpurun_out/bench_depth.json 2> gpurun_out/bench_depth.err; tail -c 1500 gpurun_out/bench_depth.json; tail -3 gpurun_out/bench_depth.err
