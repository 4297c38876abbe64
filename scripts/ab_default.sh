# A/B of environment toggles at the bench's default configuration (C2, 10 steps x 500 sweeps): usage ab_default.sh "<ENV=..>" ...
for envs in "$@"; do
env $envs timeout 400 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/ab_env.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('c2 default [$envs]', round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],3),'ms/step e2e', round(d['e2e']['value']/1e6,1), {k: round(v*1e3) for k,v in r['stage_ms_per_window'].items()}, 'launches', d['gpu_launches'])" || tail -5 gpurun_out/ab_env.err
done
