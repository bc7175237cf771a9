#!/bin/bash
# DRAM bytes moved by the kernels of ONE step of every BASELINE config (2 M pairs per GPU), from ncu's dram__bytes_{read,write}.sum.
# usage (on the GPU box): bash profiles/tools/dram_traffic.sh TAG ; then python profiles/tools/dram_traffic.py gpurun_out/TAG profiles/dram_traffic.json
TAG=${1:-r2}
mkdir -p gpurun_out
for c in 2 3 4 5; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:'seed_kernel|lanes_kernel|assemble_kernel|bin_order|class_list' -c 24 --csv --log-file gpurun_out/${TAG}_dram_cfg$c.csv \
      python bench.py --config $c --steps 1 --warmup 1 --pairs 2000000 --no-e2e --no-cpu > gpurun_out/${TAG}_dram_cfg$c.log 2>&1
  echo "dram cfg$c rc=$?"
done
