ncu --set full --clock-control none --import-source on -k regex:"k_wdedup" -s 3 -c 1 -o gpurun_out/prof_r02p_dedup -f python bench.py --steps 1 --warmup 1 --sweeps-per-step 128 --no-cpu-baseline --no-extras --groups 1 > gpurun_out/ncu_r02p_dedup.log 2>&1
ncu -i gpurun_out/prof_r02p_dedup.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_r02p_dedup_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/prof_r02p_dedup.ncu-rep gpurun_out/prof_r02p_dedup_summary.csv
