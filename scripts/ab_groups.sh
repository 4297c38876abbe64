# A/B on one box: chain groups on separate streams (the scalar kernels of one group against k_weval of another), with k_weval at 3 and at 2
# resident blocks per SM (2 leaves registers for the scalar kernels to be co-resident)
for lib in default mcmc-symreg_b200/libbsr_b200_mb2.so; do
for g in 1 2 4; do
if [ $lib = default ]; then unset BSR_LIB; else export BSR_LIB=$PWD/$lib; fi
timeout 200 python bench.py --steps 5 --warmup 3 --groups $g --no-cpu-baseline --no-extras 2>gpurun_out/ab_groups.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib groups=$g', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in d['roofline']['stage_ms_per_window'].items()})" || tail -5 gpurun_out/ab_groups.err
done
done
