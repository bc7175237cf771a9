mkdir -p gpurun_out
bash profiles/tools/gpu_round.sh r2r
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_ref.json 2> gpurun_out/r2r_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2r_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'assemble_kernel' -s 1 -c 1 -f -o gpurun_out/r2r_cfg4_prof \
    python bench.py --config 4 --steps 1 --warmup 1 --pairs 1000000 --no-e2e --no-cpu > gpurun_out/r2r_cfg4_ncu.log 2>&1; echo "ncu cfg4 rc=$?"
