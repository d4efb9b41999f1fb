# scratch driver for one gpurun call (edited per call): 2-GPU checks
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/g2_dist_test.log 2>&1; tail -n 3 gpurun_out/g2_dist_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g2_bench_2gpu.json 2> gpurun_out/g2_bench_2gpu.err; tail -c 400 gpurun_out/g2_bench_2gpu.json
