#!/bin/bash
# usage: bash tools/ncu_one.sh <name> <kernel-regex> <skip> [env...]   -> gpurun_out/ncu_<name>.ncu-rep (+ raw/source csv)
NAME=$1; REGEX=$2; SKIP=$3; shift 3
env "$@" timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$REGEX -s $SKIP -c 1 \
    -f -o gpurun_out/ncu_$NAME python tools/profile_step.py > gpurun_out/ncu_$NAME.log 2>&1
ncu -i gpurun_out/ncu_$NAME.ncu-rep --page raw --csv > gpurun_out/ncu_$NAME.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$NAME.ncu-rep --page source --csv > gpurun_out/ncu_$NAME.source.csv 2>/dev/null
ls -la gpurun_out/ncu_$NAME.*
