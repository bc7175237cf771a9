mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2s_cfg4.json 2> gpurun_out/r2s_cfg4.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r2s_cfg4.err; cut -c1-200 gpurun_out/r2s_cfg4.json
bash profiles/tools/run_sanitize.sh r2s
