mkdir -p gpurun_out
T=r2_c9
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$? $(tail -n 1 gpurun_out/${T}_tests.log)"; grep -h "FAILED\|Error" gpurun_out/${T}_tests.log | head -10
bash tools/ncu_r02.sh r02b 2>&1 | tail -12 | cut -c 1-600
bash tools/run_sanitizer.sh r02 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --graph > gpurun_out/${T}_bench_graph.json 2> gpurun_out/${T}_bench_graph.err
echo "bench graph rc=$?"; cut -c 1-250 gpurun_out/${T}_bench_graph.json; tail -n 3 gpurun_out/${T}_bench_graph.err
timeout 300 python bench.py --height 256 --width 256 --batch 1 --steps 30 --warmup 5 --no-cpu-baseline --graph > gpurun_out/${T}_256_graph.json 2> gpurun_out/${T}_256_graph.err
echo "256 graph rc=$?"; cut -c 1-250 gpurun_out/${T}_256_graph.json
timeout 300 python bench.py --height 256 --width 256 --batch 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_256.json 2> gpurun_out/${T}_256.err
echo "256 rc=$?"; cut -c 1-250 gpurun_out/${T}_256.json
