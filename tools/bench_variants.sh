#!/bin/bash
# usage: bash tools/bench_variants.sh "VAR=val VAR2=val" "..." ; prints ms/step and the slowest tc layers per variant
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/variant_$i.json 2> gpurun_out/variant_$i.err
  python - "$v" gpurun_out/variant_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(f"[{sys.argv[1]}] fps {d['value']:.2f}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.2f}  tc_ms {r['kernel_ms_per_step']:.2f}  tc TF/s {r['achieved']:.0f} frac {r['frac']:.3f}")
    for t in r["top_layers"]:
        print(f"      {t['layer']:32s} {t['ms_per_step']:7.3f} ms  {t['tflops']:6.0f} TF/s x{t['launches_per_step']:.0f}")
except Exception as e:
    print(f"[{sys.argv[1]}] FAILED {e}")
    print(open(sys.argv[2].replace('.json', '.err')).read()[-800:])
PY
done
