# length classes + merge split: GPU suite, bench configs 2 / 3 / 5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2h_pytest.log
for c in 2 3 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2h_cfg$c.json 2> gpurun_out/r2h_cfg$c.err; echo "cfg$c rc=$?"; tail -2 gpurun_out/r2h_cfg$c.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2h_cfg$c.json')); print($c, round(d['value'],1), round(d['roofline']['frac'],4), [(k['name'][:24], round(k['ms'],2)) for k in d['roofline']['kernels']])
except Exception as e: print('ERR', e)
PY
done
