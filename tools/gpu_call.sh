# scratch driver for one gpurun call (edited per call)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_conv_rsu.py -x -q -m gpu > gpurun_out/d4_rsu_test.log 2>&1
RC=$?
tail -n 3 gpurun_out/d4_rsu_test.log
if [ $RC -ne 0 ]; then echo "RSU TEST FAILED rc=$RC"; exit 0; fi
timeout 200 python tools/exp_rsu.py > gpurun_out/d4_exp_rsu.log 2>&1; tail -n 26 gpurun_out/d4_exp_rsu.log
timeout 300 python tools/ab_step.py 3 > gpurun_out/d4_ab.log 2>&1; tail -n 10 gpurun_out/d4_ab.log
