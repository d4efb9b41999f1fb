#!/bin/bash
# full-metric capture of the three streaming upfirdn2d launches (blur / upsample / downsample);  usage: bash tools/ncu_upfirdn.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upfirdn2d_stream -c 3 -f -o gpurun_out/ncu_${TAG}_upfirdn \
    python tools/upfirdn_case.py > gpurun_out/ncu_${TAG}_upfirdn.log 2>&1
ncu -i gpurun_out/ncu_${TAG}_upfirdn.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_upfirdn.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ncu_${TAG}_upfirdn.raw.csv > gpurun_out/ncu_${TAG}_upfirdn.json
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/ncu_${TAG}_upfirdn.raw.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
names=rows[h]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed_op_shared_ld.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"]
stall=[n for n in names if "issue_stalled" in n and n.endswith("_per_warp_active.pct")]
for r in rows[h+2:]:
    if len(r)!=len(names): continue
    d=dict(zip(names,r))
    print(d["Kernel Name"][:60], d.get("launch__grid_size"))
    for w in want: print("   ",w,d.get(w))
    st=sorted(((float(d[n].replace(",","")),n) for n in stall if d.get(n) not in (None,"")),reverse=True)[:6]
    for v,n in st: print("    stall",n.replace("smsp__average_warps_issue_stalled_","").replace("_per_warp_active.pct",""),v)
PY
