mkdir -p gpurun_out
free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n_pytest.log
for c in 3 5; do timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2n_cfg$c.json 2> gpurun_out/r2n_cfg$c.err; echo "cfg$c rc=$?"; tail -2 gpurun_out/r2n_cfg$c.err; cut -c1-200 gpurun_out/r2n_cfg$c.json; done
