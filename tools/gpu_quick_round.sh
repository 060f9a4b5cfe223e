set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 800 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
