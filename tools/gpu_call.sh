# scratch driver for one gpurun call (edited per call)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "fused_output_statistics or scheduling_options or instnorm_chunk" > gpurun_out/d6_new_tests.log 2>&1
RC=$?
tail -n 4 gpurun_out/d6_new_tests.log
if [ $RC -ne 0 ]; then echo "NEW TESTS FAILED rc=$RC"; grep -n "Error\|assert" gpurun_out/d6_new_tests.log | head -20; fi
(time timeout 1200 python -m pytest tests/ -x -q -m gpu) > gpurun_out/d6_pytest.log 2>&1; tail -n 8 gpurun_out/d6_pytest.log
timeout 300 python tools/ab_step.py 3 > gpurun_out/d6_ab.log 2>&1; tail -n 14 gpurun_out/d6_ab.log
