# N = 2: the bench under torchrun (headline weak scaling, sweep96 strong scaling, overlapped in-place gradient all-reduce)
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 1200 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
