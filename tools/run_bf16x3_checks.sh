#!/bin/bash
# bf16x3 (split-operand tensor-core mode) checks + bench variants. Usage on the GPU box: bash tools/run_bf16x3_checks.sh
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?  $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run bx3_conv_m0 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -x -k "bf16x3_vs_direct and -0]"
run bx3_conv_m1 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -x -k "bf16x3_vs_direct and -1]"
run bx3_shapes 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "bf16x3_work_item"
run bx3_fold 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "bf16x3_folded"
run bx3_layers 600 python -m pytest tests/test_gpu_layers.py -q -m gpu -s -k "bf16x3"
run bx3_vtoonify 600 python -m pytest tests/test_gpu_vtoonify.py -q -m gpu -s -k "bf16x3"
run bx3_psp 600 python -m pytest tests/test_gpu_psp.py -q -m gpu -s -k "bf16x3"
run bench_bx3 600 python bench.py --precision bf16x3 --steps 5 --warmup 3
run bench_tf32 600 python bench.py --steps 5 --warmup 3
cat gpurun_out/summary.txt
grep -h "err" gpurun_out/bx3_vtoonify.log gpurun_out/bx3_psp.log | head -20
