# Round-end evidence on one B200 (run under gpurun): GPU tests, smoke, bench lines, ncu launch list and full capture -> gpurun_out/
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r01j_c2.json 2> gpurun_out/bench_r01j_c2.err; tail -1 gpurun_out/bench_r01j_c2.err
python bench.py --impl reference > gpurun_out/bench_r01j_ref.json 2>/dev/null
for w in c4 c3; do python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r01j_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/launches_r01j_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu12.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wpropose|k_weval|k_wresolve|k_wclassify" -s 8 -c 4 -o gpurun_out/prof_r01j_win -f python bench.py --steps 1 --warmup 1 --sweeps-per-step 128 --no-cpu-baseline > gpurun_out/b_ncu13.log 2>&1
ls -la gpurun_out/prof_r01j_win.ncu-rep
