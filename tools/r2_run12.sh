set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest12.log; tail -6 gpurun_out/r2_pytest12.log
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 > gpurun_out/r2_modes12.jsonl 2> gpurun_out/r2_modes12.err
cat gpurun_out/r2_modes12.jsonl; tail -3 gpurun_out/r2_modes12.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_gather|roi_levels' -c 3 \
  -o gpurun_out/r2_gather_final -f python tools/infer_probe.py --images 32 --once --precision f16c8 > gpurun_out/ncu_r.log 2>&1
python tools/ncu_summary.py report gpurun_out/r2_gather_final.ncu-rep gpurun_out/r2_gather_kernels.txt
grep -E "^## launch|gpu__time_duration|smsp__inst_executed.sum|issue_active|warps_active" gpurun_out/r2_gather_kernels.txt | cut -c1-150
