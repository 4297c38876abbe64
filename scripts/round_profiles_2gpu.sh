# 2-GPU evidence (gpurun --gpus 2): multi-GPU tests, chain-sharded c2 and a row-sharded c5 slice, reference arm under torchrun
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01j_c2_2gpu.json 2> gpurun_out/b2.err; tail -2 gpurun_out/b2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --rows 2000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01j_c5_rows2M_2gpu.json 2> gpurun_out/b5.err; tail -2 gpurun_out/b5.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
