#!/bin/bash
# Run the GPU test files in separate processes (a trapped kernel poisons its CUDA context) with per-file timeouts.
# Usage (on the GPU box, from the repo root): bash tools/run_gpu_checks.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run ops 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -s
run conv_direct 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "direct_vs_torch or smalln"
for mode in 0 1; do
  run conv_tc_mode$mode 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "tc_vs_direct and -${mode}]"
done
run conv_tc_misc 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "epilogue or concat or polyphase or budget"
run conv_tc_mt 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "m_tiles"
run conv_tc_pairs 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "cta_pairs"
run conv_tc_fold 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "folded"
run conv_tc_bf16x3 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "bf16x3 or masked or tap_products or direct_global or affine"
run conv_tc_views 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "transposed_view"
run layers 900 python -m pytest tests/test_gpu_layers.py -q -m gpu -s
run vtoonify 900 python -m pytest tests/test_gpu_vtoonify.py -q -m gpu -s
run psp 600 python -m pytest tests/test_gpu_psp.py -q -m gpu -s
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
cat gpurun_out/summary.txt
