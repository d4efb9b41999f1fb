mkdir -p gpurun_out
T=r2_c2
timeout 900 python -m pytest tests/test_gpu_vtoonify.py tests/test_gpu_layers.py -q -m gpu -s > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$? $(tail -n 1 gpurun_out/${T}_tests.log)"
timeout 900 python bench.py --steps 10 --warmup 3 --dump-layers gpurun_out/${T}_layers.txt > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; cut -c 1-300 gpurun_out/${T}_bench.json; tail -n 5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --config generator --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_gen.json 2> gpurun_out/${T}_gen.err
echo "gen rc=$?"; cut -c 1-300 gpurun_out/${T}_gen.json; tail -n 3 gpurun_out/${T}_gen.err
timeout 600 python bench.py --config vtoonify_t --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_t.json 2> gpurun_out/${T}_t.err
echo "T rc=$?"; cut -c 1-300 gpurun_out/${T}_t.json; tail -n 3 gpurun_out/${T}_t.err
timeout 900 python bench.py --impl cudnn --steps 3 --warmup 1 > gpurun_out/${T}_cudnn.json 2> gpurun_out/${T}_cudnn.err
echo "cudnn rc=$?"; cut -c 1-600 gpurun_out/${T}_cudnn.json; tail -n 3 gpurun_out/${T}_cudnn.err
timeout 900 python bench.py --config video --warmup 2 > gpurun_out/${T}_video.json 2> gpurun_out/${T}_video.err
echo "video rc=$?"; cut -c 1-300 gpurun_out/${T}_video.json; tail -n 3 gpurun_out/${T}_video.err
timeout 300 python bench.py --height 256 --width 256 --batch 1 --steps 20 --warmup 3 --no-cpu-baseline --no-u8 > gpurun_out/${T}_256.json 2> gpurun_out/${T}_256.err
echo "256 rc=$?"; cut -c 1-300 gpurun_out/${T}_256.json
