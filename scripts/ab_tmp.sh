python -m pytest tests/test_gpu_window.py -x -q 2>&1 | tail -3
for w in 32 48 64; do
BSR_WINDOW=$w python bench.py --steps 6 --warmup 3 --sweeps-per-step 256 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('W=$w', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', 'windows', r['windows_profiled'], 'stages', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()})"
done
