# round 2, GPU call 4: GPU suite (attention output staging), the new bench.py (both arms)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest4.log; tail -6 gpurun_out/r2_pytest4.log
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 1500 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 600 gpurun_out/r2_bench_reference_arm.json; tail -3 gpurun_out/r2_bench_reference_arm.err
