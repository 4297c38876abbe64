for g in 1 2 3 4 6 8; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --groups $g 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('groups $g', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', 'launches', d['gpu_launches'])"
done
python -m pytest tests/test_gpu_window.py -x -q 2>&1 | tail -2
