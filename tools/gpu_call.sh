mkdir -p gpurun_out
T=r2_c7
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_conv_rs.py -x -q -m gpu 2>&1 | tail -n 3; done
timeout 900 python tools/diag_rs.py > gpurun_out/${T}_diag.log 2>&1; grep "B[14]" gpurun_out/${T}_diag.log | cut -c 1-200
timeout 600 python tools/diag_batch.py 2>&1 | tail -4
timeout 600 python tools/tc_role_timing.py bf16x3 rs 2>&1 | grep -A3 "plain\|rs_cg 2" | cut -c 1-330 > gpurun_out/${T}_roles.log; cat gpurun_out/${T}_roles.log
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_layers.py tests/test_gpu_vtoonify.py -q -m gpu 2>&1 | tail -n 3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-u8 --dump-layers gpurun_out/${T}_layers.txt > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; cut -c 1-200 gpurun_out/${T}_bench.json; head -12 gpurun_out/${T}_layers.txt
VT_TC_STRICT=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-u8 > gpurun_out/${T}_bench_strict.json 2> gpurun_out/${T}_bench_strict.err
echo "strict bench:"; cut -c 1-200 gpurun_out/${T}_bench_strict.json
