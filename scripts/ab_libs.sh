# A/B of library variants (scripts/build_variant.sh) on one box: usage ab_libs.sh <workload> <suffix|default> ...
w=$1; shift
for v in "$@"; do
if [ $v = default ]; then unset BSR_LIB; else export BSR_LIB=$PWD/mcmc-symreg_b200/libbsr_b200_$v.so; fi
sw=""; [ $w = c2 ] && sw="--steps 5 --sweeps-per-step 256"; [ $w != c2 ] && sw="--steps 2"
timeout 400 python bench.py --workload $w --warmup 3 $sw --no-cpu-baseline --no-extras 2>gpurun_out/ab_lib.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w lib=$v', round(d['value']/1e6,4),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()})" || tail -5 gpurun_out/ab_lib.err
done
