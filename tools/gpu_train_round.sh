#!/usr/bin/env bash
# One GPU visit for the training headline: parity tests, bench (both arms), ncu launch list of the training step and
# full captures of its top kernels, summarised on the box (the raw csv / reports can exceed the copy-back limit).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r1t}
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/${TAG}_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -n 5 gpurun_out/${TAG}_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2>/dev/null; echo "rc=$?"
if [ "${SKIP_NCU:-0}" != 1 ]; then
echo "== ncu launch list (training step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file /tmp/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-inference --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.out 2>&1; echo "rc=$?"
python tools/ncu_summary.py launches /tmp/launches.csv gpurun_out/${TAG}_launches.txt; head -n 14 gpurun_out/${TAG}_launches.txt
echo "== ncu full: training kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn2_kernel|gemm_tc2_kernel|attention_bwd|ln_bwd_kernel|attention_mma" -s 60 -c 24 -f -o /tmp/prof_train \
   python bench.py --steps 1 --warmup 1 --no-inference --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.out 2>&1; echo "rc=$?"
python tools/ncu_summary.py report /tmp/prof_train.ncu-rep gpurun_out/${TAG}_train_ncu.txt
ls -la /tmp/prof_train.ncu-rep; [ $(stat -c %s /tmp/prof_train.ncu-rep) -lt 40000000 ] && cp /tmp/prof_train.ncu-rep gpurun_out/${TAG}_train.ncu-rep
fi
ls -la gpurun_out | tail -n 12
