mkdir -p gpurun_out
T=r2_c14
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -n 2
timeout 600 python tools/upfirdn_bench.py > gpurun_out/${T}_upfirdn.log 2>&1; cat gpurun_out/${T}_upfirdn.log
timeout 600 python tools/smalln_bench.py > gpurun_out/${T}_smalln.log 2>&1; tail -n 10 gpurun_out/${T}_smalln.log
