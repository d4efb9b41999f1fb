# scratch driver for one gpurun call (edited per call)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_conv_rsu.py -x -q -m gpu -s > gpurun_out/d2_rsu_test.log 2>&1
RC=$?
tail -n 4 gpurun_out/d2_rsu_test.log
if [ $RC -ne 0 ]; then echo "RSU EPI1 FAILED rc=$RC -> VT_RSU_EPI=0"; export VT_RSU_EPI=0; fi
timeout 200 python tools/exp_rsu.py > gpurun_out/d2_exp_rsu.log 2>&1; tail -n 12 gpurun_out/d2_exp_rsu.log
(time timeout 1200 python -m pytest tests/ -x -q -m gpu) > gpurun_out/d2_pytest.log 2>&1; tail -n 8 gpurun_out/d2_pytest.log
python bench.py > gpurun_out/d2_bench.json 2> gpurun_out/d2_bench.err; tail -c 300 gpurun_out/d2_bench.json
VT_INSTNORM_CHUNKS=0 python bench.py --steps 10 > gpurun_out/d2_bench_chunks0.json 2> gpurun_out/d2_bench_chunks0.err
