set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 600 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_reference_arm_n$N.json 2> gpurun_out/r2_bench_reference_arm_n$N.err
tail -c 300 gpurun_out/r2_bench_reference_arm_n$N.json
