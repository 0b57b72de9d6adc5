#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run4_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run4_pytest.txt
tail -15 gpurun_out/r2_run4_pytest.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_run4_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_run4_bench.json') if l.startswith('{')][-1]
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],2), round(v.get('frac',0),3))
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('e2e'), v.get('gpu_launches'), v.get('error'), {a:b['ms_per_step'] for a,b in (v.get('kernels') or {}).items()})
PY
