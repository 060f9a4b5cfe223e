set -x
mkdir -p gpurun_out
VETO_TRAIN_RECOMPUTE=1 timeout 1200 python -m pytest tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -4
for v in 0 1 0 1; do
VETO_TRAIN_RECOMPUTE=$v timeout 600 python bench.py --legs train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b31_$v.json 2> gpurun_out/b31_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/b31_$v.json'))
t=d.get('train') or d
print('recompute=$v', t.get('value'), t.get('ms_per_step'), t.get('train_workspace_gb'), t.get('config',{}).get('workload','')[:40])
PY
done
