mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r2t_smoke.log
bash profiles/tools/gpu_round.sh r2t
