# End-of-round confirmation on one B200: all GPU tests, smoke, both bench arms (what the driver runs)
set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 300 python tools/depth_bench.py 2>&1 | tail -1 | tee gpurun_out/depth_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -c 300 gpurun_out/bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
