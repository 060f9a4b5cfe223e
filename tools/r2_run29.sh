set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes29.*
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
for d in 0 0; do
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes29.jsonl 2>> gpurun_out/r2_modes29.err &
sleep 25; nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader; sleep 2; nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
wait
done
cat gpurun_out/r2_modes29.jsonl; tail -5 gpurun_out/r2_modes29.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/b29.json 2> gpurun_out/b29.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/b29.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'])
PY
