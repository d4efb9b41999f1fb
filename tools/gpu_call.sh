mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -n 1; done
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:fir4_nhwc -c 3 -f -o gpurun_out/ncu_r2c23_fir python tools/profile_step.py > gpurun_out/ncu_r2c23_fir.log 2>&1
ncu -i gpurun_out/ncu_r2c23_fir.ncu-rep --page raw --csv > gpurun_out/ncu_r2c23_fir.raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/ncu_r2c23_fir.raw.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
names=rows[h]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread",
 "l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct"]
for r in rows[h+2:]:
    if len(r)!=len(names): continue
    d=dict(zip(names,r))
    print(d["Kernel Name"][:50], d.get("launch__grid_size"))
    for w in want: print("   ",w,d.get(w))
PY
