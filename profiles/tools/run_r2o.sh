mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json')); print(round(d['value'],1), [(k['name'][:20], round(k['ms'],3)) for k in d['roofline']['kernels']])
PY
