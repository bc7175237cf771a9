mkdir -p gpurun_out
bash profiles/tools/dram_traffic.sh r2m
python profiles/tools/dram_traffic.py gpurun_out/r2m gpurun_out/r2m_dram_traffic.json
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2m_cfg$c.json 2> gpurun_out/r2m_cfg$c.err; echo "cfg$c rc=$?"; cut -c1-200 gpurun_out/r2m_cfg$c.json; done
