mkdir -p gpurun_out
timeout 120 python tools/exp_rsu256.py > gpurun_out/d7_rsu256.log 2>&1; tail -n 6 gpurun_out/d7_rsu256.log
