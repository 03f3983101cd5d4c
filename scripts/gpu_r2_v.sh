mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 --timeout-method=thread > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2v_pytest.log
echo "== probe"; timeout 600 python scripts/scan_tail_probe.py --rows 10000,1250000,10000000 > gpurun_out/r2v_probe.jsonl 2> gpurun_out/r2v_probe.err; python - <<'PY'
import json
for l in open('gpurun_out/r2v_probe.jsonl'):
    r=json.loads(l)
    if 'storage' in r:
        print(r['rows'], r['options'][:40].ljust(40), r['ms_per_search'], r.get('scan_loop_us'), r.get('cta_end_spread_us'), r.get('epilogue_us'), r.get('fused_finalize_us'), r.get('last_cta_us'))
PY
