# A/B of one build on one box: window / parity GPU tests, then c2 / c4 / c3 / c5 throughput and stage times (LABEL names the variant)
LABEL=${LABEL:-variant}
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for w in c2 c4 c3 c5; do
st=5; sw=""; [ $w = c2 ] && sw="--sweeps-per-step 256"; [ $w != c2 ] && st=2
timeout 400 python bench.py --workload $w --steps $st --warmup 3 $sw --no-cpu-baseline --no-extras 2>gpurun_out/ab_${LABEL}_$w.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$LABEL $w', round(d['value']/1e6,4),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()}, {k: round(v*1e3) for k,v in r['kernel_ms'].items()}, 'exec/ref', round(d['node_evals_exec_per_sec']/d['node_evals_ref_per_sec'],3), 'wide', d['fp64_sweeps'], 'acc', round(d['accept_rate'],5), 'rankrej', round(d['rank_reject_rate'],4))" || tail -5 gpurun_out/ab_${LABEL}_$w.err
done
