set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_modes17.* gpurun_out/r2_pytest17*.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_pytest17.log
cat gpurun_out/r2_pytest17.log
for r in 1 2; do
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision f16c8 >> gpurun_out/r2_modes17.jsonl 2>> gpurun_out/r2_modes17.err
done
timeout 300 python tools/infer_probe.py --images 32 --chunks 0 --steps 3 --warmup 2 --precision bf16x3 >> gpurun_out/r2_modes17.jsonl 2>> gpurun_out/r2_modes17.err
cat gpurun_out/r2_modes17.jsonl; tail -5 gpurun_out/r2_modes17.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/b17.json 2> gpurun_out/b17.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/b17.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('legs',{}).items()} if 'legs' in d else [k for k in d])
PY
