# A/B of environment toggles (tile rows, row splits) on one box: usage ab_env.sh <workload> "<ENV=.. ENV=..>" ...
w=$1; shift
for envs in "$@"; do
env $envs timeout 400 python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline --no-extras 2>gpurun_out/ab_env.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w [$envs]', round(d['value']/1e6,4),'M/s', round(d['ms_per_step'],2),'ms/step', {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()})" || tail -5 gpurun_out/ab_env.err
done
