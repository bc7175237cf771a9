mkdir -p gpurun_out
bash profiles/tools/gpu_round.sh r2l
