python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --sweeps-per-step 256 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', 'stages', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()})"
