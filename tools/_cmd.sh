mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],l['e2e']['seconds_per_call'],'cores',(l['cpu_baseline'] or {}).get('cores'), 'cpu',(l['cpu_baseline'] or {}).get('value'))
print('fast',l['fast']['value'],l['fast']['ms_per_step'],l['fast']['e2e']['value'], l['roofline_fast']['frac'], l['fast']['validation_vs_exact'])
print('cfg5',l['nhood_cfg5_strong'])
m=l['moran']; print('moran',m.get('value'),m.get('ms_per_step'),m.get('e2e',{}).get('seconds_all'), m.get('n_perms_100_seconds'), m.get('roofline',{}).get('frac'), m.get('error'))
print('cooc',l['co_occurrence'].get('e2e',{}).get('seconds_all'), l['co_occurrence'].get('kernel_ms'), l['co_occurrence'].get('error'))
print('rip',l['ripley_L'].get('e2e',{}).get('seconds_all'), l['ripley_L'].get('kernel_ms'), l['ripley_L'].get('error'))
PY
