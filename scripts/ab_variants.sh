# A/B of one build on one box: GPU tests, then c2 / c4 / c3 throughput and stage times (edit the variant label / env toggles)
timeout 900 python -m pytest tests/test_gpu_window.py -q -m gpu -x 2>&1 | tail -3
for v in hoist; do
for w in c2 c4 c3; do
st=5; sw=""; [ $w = c2 ] && sw="--sweeps-per-step 256"; [ $w != c2 ] && st=3
timeout 300 python bench.py --workload $w --steps $st --warmup 3 $sw --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v $w', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()}, {k: round(v*1e3) for k,v in r['kernel_ms'].items()}, 'exec/ref', round(d['node_evals_exec_per_sec']/d['node_evals_ref_per_sec'],3))"
done
done
