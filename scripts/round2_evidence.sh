# Round-2 evidence on one B200 (run under gpurun): GPU tests, smoke, the bench lines (ours + reference arm), ncu launch list and one
# `--set full` capture of the window kernels -> gpurun_out/.  TAG names the files; profiles/ is filled from them afterwards.
TAG=${TAG:-r02r}
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-400
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -1 gpurun_out/bench_${TAG}.err | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/launches_${TAG}_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wpropose|k_weval|k_wresolve|k_wclassify|k_wdedup" -s 10 -c 5 -o gpurun_out/prof_${TAG}_win -f python bench.py --steps 1 --warmup 1 --sweeps-per-step 128 --groups 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_full.log 2>&1
ls -la gpurun_out/prof_${TAG}_win.ncu-rep
ncu -i gpurun_out/prof_${TAG}_win.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_win.ncu-rep gpurun_out/prof_${TAG}_summary.csv
cut -c1-200 gpurun_out/bench_${TAG}.json
# rank / decision census against the oracle (2 x 1e5 proposals) and one full capture of k_weval on a slice of the large-data workload
python scripts/census_parity.py --shapes c2,c4 --out gpurun_out/census_${TAG}.json > gpurun_out/census_${TAG}.log 2>&1; grep -c . gpurun_out/census_${TAG}.log
ncu --set full --clock-control none -k regex:"k_weval" -s 1 -c 1 -o gpurun_out/prof_${TAG}_c5 -f python bench.py --workload c5 --rows 4000000 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_c5.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_c5.ncu-rep gpurun_out/prof_${TAG}_c5_summary.csv
