mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
for f in test_gpu_bisenet test_gpu_gradfix test_gpu_frame_loop test_gpu_dist test_gpu_fullsize; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -s > gpurun_out/r2_c1_$f.log 2>&1
  echo "$f rc=$? $(tail -n 1 gpurun_out/r2_c1_$f.log)"
done
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_bisenet.py --deselect tests/test_gpu_gradfix.py --deselect tests/test_gpu_frame_loop.py --deselect tests/test_gpu_dist.py --deselect tests/test_gpu_fullsize.py > gpurun_out/r2_c1_rest.log 2>&1
echo "rest rc=$? $(tail -n 1 gpurun_out/r2_c1_rest.log)"
timeout 600 python tools/tc_role_timing.py > gpurun_out/r2_c1_roles.log 2>&1
tail -n 25 gpurun_out/r2_c1_roles.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-layers gpurun_out/r2_c1_layers.txt > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
echo "bench rc=$?"; cut -c 1-400 gpurun_out/r2_c1_bench.json
