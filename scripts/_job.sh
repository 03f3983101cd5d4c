mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3),'cpu',round(d['cpu_baseline']['value'],2),d['clocks'])
for c in d['configs']: print(c['config'], round(c['ms_per_search'],4), round(c['queries_per_s']), c['roofline']['bound'], round(c['roofline']['frac'],3), round(c['roofline'].get('frac_whole_search',0),3))
PY
