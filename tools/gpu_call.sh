bash tools/ncu_r02.sh r2c28 2>&1 | tail -n 9
