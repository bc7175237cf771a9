mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2i_pytest.log
timeout 900 python bench.py --path api > gpurun_out/r2i_api.json 2> gpurun_out/r2i_api.err; echo "api rc=$?"; tail -3 gpurun_out/r2i_api.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_api.json'))
for k in ('ours','reference'):
    print(k,[(r['threads'],round(r['mpairs_per_s'],3)) for r in d['by_threads'][k]])
PY
lscpu | grep -E "^CPU\(s\)|Socket|NUMA|Model name|Thread" ; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nvidia-smi topo -m 2>/dev/null | head -20
