# Round-2 profile evidence on one B200 (run under gpurun): ncu launch list of the bench command and one `--set full` capture of the four
# window kernels (source view included) -> gpurun_out/.  TAG names the files.
TAG=${TAG:-r02a}
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/launches_${TAG}_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wpropose|k_weval|k_wresolve|k_wclassify" -s 8 -c 4 -o gpurun_out/prof_${TAG}_win -f python bench.py --steps 1 --warmup 1 --sweeps-per-step 128 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_full.log 2>&1
ls -la gpurun_out/prof_${TAG}_win.ncu-rep
ncu -i gpurun_out/prof_${TAG}_win.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_win.ncu-rep gpurun_out/prof_${TAG}_summary.csv
