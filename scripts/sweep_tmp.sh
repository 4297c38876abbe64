python -m pytest tests/test_gpu_window.py -x -q 2>&1 | tail -4
for g in 1 2 4 8; do for s in 100 300; do
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --groups $g --sweeps-per-step $s 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('groups',$g,'S',$s, round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step launches', d['gpu_launches'])"
done; done
