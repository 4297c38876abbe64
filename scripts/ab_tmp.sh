python -m pytest tests/test_gpu_window.py -q -x 2>&1 | tail -2
for co in default 100 75 68 50; do
export BSR_WEVAL_CARVEOUT=$co; [ $co = default ] && unset BSR_WEVAL_CARVEOUT
python bench.py --steps 5 --warmup 3 --sweeps-per-step 256 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('carveout $co', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in r['kernel_ms'].items()})"
done
