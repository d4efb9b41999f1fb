#!/bin/bash
# Round-2 profile set (GPU box): launch list of one VToonify-D step, full-metric captures of the row-strip launches (64->64 and
# 32->32 conv2 of the last two levels) and of the median conv_tc launch (512->512 @72x128).   usage: bash tools/ncu_r02.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launch_summary_${TAG}.txt
head -24 gpurun_out/launch_summary_${TAG}.txt
MEDIAN=$(python - <<PY
import csv
lines=[l for l in open("gpurun_out/launches_${TAG}.csv") if not l.startswith("==")]
d=[float(r["Metric Value"].replace(",","")) for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum" and "conv_tc" in r["Kernel Name"]]
order=sorted(range(len(d)), key=lambda i: d[i])
print(order[len(order)//2])
PY
)
echo "conv_tc median launch index: $MEDIAN"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_rs_kernel -c 3 \
    -f -o gpurun_out/ncu_${TAG}_rs python tools/profile_step.py > gpurun_out/ncu_${TAG}_rs.log 2>&1
ncu -i gpurun_out/ncu_${TAG}_rs.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_rs.raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_rsu_kernel -c 3 \
    -f -o gpurun_out/ncu_${TAG}_rsu python tools/profile_step.py > gpurun_out/ncu_${TAG}_rsu.log 2>&1
ncu -i gpurun_out/ncu_${TAG}_rsu.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_rsu.raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s $MEDIAN -c 1 \
    -f -o gpurun_out/ncu_${TAG}_median python tools/profile_step.py > gpurun_out/ncu_${TAG}_median.log 2>&1
ncu -i gpurun_out/ncu_${TAG}_median.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_median.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ncu_${TAG}_rs.raw.csv gpurun_out/ncu_${TAG}_rsu.raw.csv gpurun_out/ncu_${TAG}_median.raw.csv > gpurun_out/ncu_full_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/ncu_full_${TAG}.json"))
for k,v in d.items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("kernel","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","launch__grid_size","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread")})
PY
