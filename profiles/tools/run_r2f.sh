mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --path api > gpurun_out/r2f_api.json 2> gpurun_out/r2f_api.err; echo "api rc=$?"; tail -3 gpurun_out/r2f_api.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_api.json'))
for k in ('ours','reference'):
    print(k,[(r['threads'],round(r['mpairs_per_s'],3)) for r in d['by_threads'][k]])
PY
