mkdir -p gpurun_out
timeout 600 python bench.py --per-base-p --steps 5 --warmup 3 --no-cpu > gpurun_out/r2z_bench_p.json 2> gpurun_out/r2z_bench_p.err; echo "rc=$?"; tail -3 gpurun_out/r2z_bench_p.err; cut -c1-300 gpurun_out/r2z_bench_p.json
