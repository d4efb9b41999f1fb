mkdir -p gpurun_out
T=r2_c12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
echo "bench N=2 rc=$?"; cut -c 1-300 gpurun_out/${T}_bench2.json; tail -n 5 gpurun_out/${T}_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config video --warmup 2 > gpurun_out/${T}_video2.json 2> gpurun_out/${T}_video2.err
echo "video N=2 rc=$?"; cut -c 1-300 gpurun_out/${T}_video2.json; tail -n 5 gpurun_out/${T}_video2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config video --wire f32 --warmup 2 > gpurun_out/${T}_video2_f32.json 2> gpurun_out/${T}_video2_f32.err
echo "video f32 N=2 rc=$?"; cut -c 1-300 gpurun_out/${T}_video2_f32.json; tail -n 3 gpurun_out/${T}_video2_f32.err
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -n 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_ref2.json 2> gpurun_out/${T}_ref2.err
echo "ref N=2 rc=$?"; cut -c 1-300 gpurun_out/${T}_ref2.json
