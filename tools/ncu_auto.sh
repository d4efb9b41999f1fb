#!/bin/bash
# Launch list of one VToonify-D step + full-metric captures of (a) the longest conv_tc launch and (b) the median one
# (the 512->512 @72x128 layers dominate the launch count).  usage (GPU box): bash tools/ncu_auto.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launch_summary_${TAG}.txt
head -16 gpurun_out/launch_summary_${TAG}.txt
read LONGEST MEDIAN <<< $(python - <<PY
import csv
lines=[l for l in open("gpurun_out/launches_${TAG}.csv") if not l.startswith("==")]
d=[float(r["Metric Value"].replace(",","")) for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum" and "conv_tc" in r["Kernel Name"]]
order=sorted(range(len(d)), key=lambda i: d[i])
print(order[-1], order[len(order)//2])
PY
)
echo "conv_tc launch indices: longest=$LONGEST median=$MEDIAN"
for pair in "long:$LONGEST" "median:$MEDIAN"; do
  NAME=${pair%%:*}; SKIP=${pair##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s $SKIP -c 1 \
      -f -o gpurun_out/ncu_${TAG}_$NAME python tools/profile_step.py > gpurun_out/ncu_${TAG}_$NAME.log 2>&1
  ncu -i gpurun_out/ncu_${TAG}_$NAME.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_$NAME.raw.csv 2>/dev/null
done
python tools/ncu_summary.py gpurun_out/ncu_${TAG}_long.raw.csv gpurun_out/ncu_${TAG}_median.raw.csv > gpurun_out/ncu_full_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/ncu_full_${TAG}.json"))
for k,v in d.items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("kernel","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__grid_size","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread")})
PY
