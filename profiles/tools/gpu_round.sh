#!/bin/bash
# One gpurun call of a development round: GPU parity suite, bench (N = 1), ncu launch list, ncu full capture of the kernels of a step.
# usage: gpu_round.sh TAG [pytest-args...]     outputs under gpurun_out/TAG_*
TAG=${1:-r2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'seed_kernel|lanes_kernel|assemble_kernel|bin_order|pack_kernel|class_list' -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --pairs 2000000 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_seed_kernel|assemble_lanes_kernel' -s 2 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 1 --pairs 2000000 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
