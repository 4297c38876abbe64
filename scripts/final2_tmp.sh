timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/b2.err | tail -1 > gpurun_out/bench_r01i_c2_2gpu.json
cut -c1-200 gpurun_out/bench_r01i_c2_2gpu.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload c5 --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/b5.err | tail -1 > gpurun_out/bench_r01i_c5_rows2M_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r01i_c5_rows2M_2gpu.json')); r=d['roofline']; print('c5 x2', d['value'], 'props/s', d['ms_per_step'], 'ms/step', 'node evals exec/s %.3g ref %.3g' % (d['node_evals_exec_per_sec'], d['node_evals_ref_per_sec']), r['stage_ms_per_window'], r['kernel_ms'])"
