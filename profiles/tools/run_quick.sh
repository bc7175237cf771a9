# quick check of a lane/sweep kernel change: lane + full-size + golden GPU tests, kernel-only bench of configs 2 and 5
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for c in 2 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_cfg$c.json 2> gpurun_out/${TAG}_cfg$c.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_cfg$c.json')); print(round(d['value'],1), [(k['name'][:20], round(k['ms'],3)) for k in d['roofline']['kernels']], d['gpu_launches'])
PY
done
