mkdir -p gpurun_out
T=r2_c26
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_4gpu.json 2> gpurun_out/${T}_bench_4gpu.err; tail -c 1500 gpurun_out/${T}_bench_4gpu.json; tail -n 5 gpurun_out/${T}_bench_4gpu.err
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -n 1
