#!/bin/bash
# Full-metric ncu captures of selected launches of one VToonify-D step (tools/profile_step.py window).
# usage: bash tools/ncu_capture.sh <tag>   -> gpurun_out/ncu_<tag>_*.ncu-rep + raw csv
TAG=${1:-r01}
mkdir -p gpurun_out
cap() { # name kernel-regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 \
      -f -o gpurun_out/ncu_${TAG}_$1 python tools/profile_step.py > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/ncu_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_$1.raw.csv 2>/dev/null
}
cap tc_512 conv_tc 10 1       # 512->512 3x3 @72x128 (encoder res block)
cap tc_256 conv_tc 48 1
cap tc_last conv_tc 60 1      # 32->32 3x3 @2304x4096 (convs.15)
cap fir_last fir_nhwc 4 1
cap torgb_last smalln_conv_kernel.3 9 1
ls -la gpurun_out/*.ncu-rep
