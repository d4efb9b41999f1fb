# scratch driver for one gpurun call (edited per call)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s > gpurun_out/f3_fullsize.log 2>&1; grep -i "err\|passed\|failed" gpurun_out/f3_fullsize.log | tail -n 12
python bench.py --dump-layers gpurun_out/f3_layers.txt > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; tail -c 200 gpurun_out/f3_bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/f3_launches.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/f3_launches.csv > gpurun_out/f3_launch_summary.txt; head -18 gpurun_out/f3_launch_summary.txt
bash tools/ncu_one.sh f3_tc conv_tc_kernel 12 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; tail -n 3 gpurun_out/f3_smoke.log
