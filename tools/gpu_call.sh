# scratch driver for one gpurun call (edited per call): 2-GPU bench with the final code
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g3_bench_2gpu.json 2> gpurun_out/g3_bench_2gpu.err; tail -c 300 gpurun_out/g3_bench_2gpu.json; tail -n 3 gpurun_out/g3_bench_2gpu.err
