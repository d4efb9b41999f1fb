"""Summarise `ncu --page raw --csv` exports into one JSON/markdown block (the numbers quoted in profiles/ and DESIGN.md)."""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_gmma.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    if hdr is None:
        return []
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {n: (v, u) for n, v, u in zip(names, r, units)}
        out.append(d)
    return out


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return s


if __name__ == "__main__":
    res = {}
    for path in sys.argv[1:]:
        for idx, d in enumerate(load(path)):
            item = {"kernel": d["Kernel Name"][0].split("(")[0]}
            for k, (v, u) in d.items():
                if any(k == key or k.startswith(key) for key in KEYS) or "tensor" in k:
                    item[k] = [num(v), u]
            res[path.split("/")[-1] + (f"#{idx}" if idx else "")] = item
    print(json.dumps(res, indent=1))
