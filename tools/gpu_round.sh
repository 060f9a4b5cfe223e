#!/usr/bin/env bash
# One GPU visit: parity tests, bench, chunk sweep, ncu launch list + full captures of the top kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${SKIP_SWEEP:-0}" != 1 ]; then
echo "== sweep"; DIAG_IMAGES=8 DIAG_CHUNKS=${DIAG_CHUNKS:-2048,4096,8192} timeout 600 python tests/gpu_diag.py speed:bf16x3 speed:bf16 2>&1 | grep '^{'
fi
if [ "${SKIP_NCU:-0}" != 1 ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --images 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.out 2>&1; echo "rc=$?"
echo "== ncu full: gemm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 30 -c 8 -f -o gpurun_out/prof_gemm \
   python bench.py --images 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.out 2>&1; echo "rc=$?"
echo "== ncu full: attention + layernorm + tokens + gather"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_mma_kernel|layernorm_kernel|tokens_kernel|roi_gather_kernel|gemm_simt" -s 4 -c 8 -f -o gpurun_out/prof_rows \
   python bench.py --images 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_rows.out 2>&1; echo "rc=$?"
fi
ls -la gpurun_out | head -30
