# round 2 record run (final build) on one B200: GPU suite, smoke, both bench arms, ncu launch list + full captures summarised on the box
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
bash tools/r2_final2.sh
