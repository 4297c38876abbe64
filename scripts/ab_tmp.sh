for v in base p12 p16 r8; do
lib=$PWD/mcmc-symreg_b200/libbsr_b200.so; [ $v != base ] && lib=$PWD/mcmc-symreg_b200/libbsr_b200_$v.so
BSR_LIB=$lib python bench.py --steps 5 --warmup 3 --sweeps-per-step 256 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()})"
done
