mkdir -p gpurun_out
T=r2_c8
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$? $(tail -n 1 gpurun_out/${T}_tests.log)"; grep -h "FAILED\|Error" gpurun_out/${T}_tests.log | head -10
timeout 900 python bench.py --impl cudnn --steps 3 --warmup 1 > gpurun_out/${T}_cudnn.json 2> gpurun_out/${T}_cudnn.err
echo "cudnn rc=$?"; cut -c 1-700 gpurun_out/${T}_cudnn.json; tail -n 3 gpurun_out/${T}_cudnn.err
timeout 900 python bench.py --impl cudnn --config generator --steps 5 --warmup 2 > gpurun_out/${T}_cudnn_gen.json 2> gpurun_out/${T}_cudnn_gen.err
echo "cudnn gen rc=$?"; cut -c 1-400 gpurun_out/${T}_cudnn_gen.json
bash tools/ncu_r02.sh r02a 2>&1 | tail -40
