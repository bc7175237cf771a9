mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
timeout 900 python bench.py --path api > gpurun_out/r2e_api.json 2> gpurun_out/r2e_api.err; echo "api rc=$?"; tail -3 gpurun_out/r2e_api.err; cut -c1-600 gpurun_out/r2e_api.json
timeout 300 python bench.py --path copy > gpurun_out/r2e_copy.json 2> gpurun_out/r2e_copy.err; echo "copy rc=$?"; cat gpurun_out/r2e_copy.json
