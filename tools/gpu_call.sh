mkdir -p gpurun_out
T=r2_c15
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv_rs.py tests/test_gpu_layers.py tests/test_gpu_vtoonify.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -n 5
timeout 600 python tools/smalln_bench.py > gpurun_out/${T}_smalln.log 2>&1; tail -n 12 gpurun_out/${T}_smalln.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench.json; tail -n 3 gpurun_out/${T}_bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv | head -n 24
