# Last evidence set of the round (one gpurun call): window / parity tests, same-box A/B against the library of c5436f2, bench lines,
# ncu launch list + full capture, census.  TAG names the files.
TAG=${TAG:-r02z}
timeout 600 python -m pytest tests/test_gpu_window.py tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_experiments.py -q -m gpu -x 2>&1 | tail -2
[ -f mcmc-symreg_b200/libbsr_b200_r02v.so ] && { bash scripts/ab_libs.sh c4 r02v default; bash scripts/ab_libs.sh c5 r02v default; bash scripts/ab_libs.sh c3 r02v default; }
unset BSR_LIB
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -1 gpurun_out/bench_${TAG}.err | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/launches_${TAG}_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wpropose|k_weval|k_wresolve|k_wclassify|k_wdedup" -s 10 -c 5 -o gpurun_out/prof_${TAG}_win -f python bench.py --steps 1 --warmup 1 --sweeps-per-step 128 --groups 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_win.ncu-rep gpurun_out/prof_${TAG}_summary.csv
python scripts/census_parity.py --shapes c2,c4 --out gpurun_out/census_${TAG}.json > gpurun_out/census_${TAG}.log 2>&1; grep -c . gpurun_out/census_${TAG}.log
cut -c1-160 gpurun_out/bench_${TAG}.json
