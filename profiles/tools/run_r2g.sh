# round-2 re-anchor: GPU parity suite, N=1 bench, reference arm, api + copy benches, ncu launch list + full capture
mkdir -p gpurun_out
bash profiles/tools/gpu_round.sh r2g
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2g_ref.json
timeout 900 python bench.py --path api > gpurun_out/r2g_api.json 2> gpurun_out/r2g_api.err; echo "api rc=$?"; tail -3 gpurun_out/r2g_api.err; cut -c1-1200 gpurun_out/r2g_api.json
timeout 300 python bench.py --path copy > gpurun_out/r2g_copy.json 2> gpurun_out/r2g_copy.err; echo "copy rc=$?"; cat gpurun_out/r2g_copy.json
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/r2g_cfg$c.json 2> gpurun_out/r2g_cfg$c.err; echo "cfg$c rc=$?"; cut -c1-400 gpurun_out/r2g_cfg$c.json; done
