mkdir -p gpurun_out
bash profiles/tools/gpu_round.sh r2w
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2w_cfg3.json 2> gpurun_out/r2w_cfg3.err; echo "cfg3 rc=$?"; cut -c1-160 gpurun_out/r2w_cfg3.json
