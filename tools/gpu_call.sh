mkdir -p gpurun_out
T=r2_c19
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backward.py -q -m gpu -x 2>&1 | tail -n 8
timeout 600 python tools/upfirdn_bench.py > gpurun_out/${T}_upfirdn.log 2>&1; cat gpurun_out/${T}_upfirdn.log
bash tools/ncu_upfirdn.sh r2c19
