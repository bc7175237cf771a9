#!/bin/bash
# One multi-GPU gpurun call (usage: gpurun --gpus N -- 'bash profiles/tools/run_multi.sh N TAG'): the link's ceiling with every rank copying
# at once, BASELINE config 5 (100 M mixed-length pairs over 8 GPUs: 12.5 M per GPU) and config 2 sharded over N ranks, and the
# reference-shaped API (panda_run_pool) spreading its workers over the N GPUs.
N=${1:-8}; TAG=${2:-r2n$N}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1; lscpu | grep -E "^CPU\(s\)|Socket|NUMA|Model name|Thread" >> gpurun_out/${TAG}_topo.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --path copy --steps 3 > gpurun_out/${TAG}_copy.json 2> gpurun_out/${TAG}_copy.err; echo "copy rc=$?"; cut -c1-900 gpurun_out/${TAG}_copy.json
timeout 900 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/${TAG}_cfg5.err; cut -c1-300 gpurun_out/${TAG}_cfg5.json
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_cfg2.json 2> gpurun_out/${TAG}_cfg2.err; echo "cfg2 rc=$?"; tail -2 gpurun_out/${TAG}_cfg2.err; cut -c1-300 gpurun_out/${TAG}_cfg2.json
for t in 4 8 16 32; do timeout 300 pandaseq_b200/api_bench 4000000 150 $t > gpurun_out/${TAG}_api_t$t.json 2> gpurun_out/${TAG}_api_t$t.err; echo "api t=$t rc=$?"; cat gpurun_out/${TAG}_api_t$t.json; done
PANDASEQ_B200_DEVICES=1 timeout 300 pandaseq_b200/api_bench 4000000 150 8 > gpurun_out/${TAG}_api_1gpu_t8.json 2>&1; cat gpurun_out/${TAG}_api_1gpu_t8.json
