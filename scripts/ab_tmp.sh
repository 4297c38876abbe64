

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload c5 --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/b5.err | tail -1 > gpurun_out/bench_r01g_c5_2gpu.json
tail -3 gpurun_out/b5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01g_c5_2gpu.json')); r=d['roofline']; print('c5 x2', d['value'], 'props/s', d['ms_per_step'], 'ms/step', 'node evals exec/s %.3g ref %.3g' % (d['node_evals_exec_per_sec'], d['node_evals_ref_per_sec']), r['stage_ms_per_window'], r['kernel_ms'], d['config'])"
