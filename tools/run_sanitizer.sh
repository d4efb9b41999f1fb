#!/bin/bash
# compute-sanitizer passes over tools/sanitize_case.py (GPU box).  memcheck: out-of-bounds / misaligned accesses of every kernel;
# racecheck: shared-memory hazards of the warp-specialised pipelines (hand-rolled mbarrier protocols: barriers are reported as
# hazards only when a generic-proxy access really races).  usage: bash tools/run_sanitizer.sh <tag> [case file] [timeout s]
TAG=${1:-r02}
CASE=${2:-tools/sanitize_case.py}
TO=${3:-1200}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout $TO compute-sanitizer --tool $tool --print-limit 20 python $CASE > gpurun_out/sanitizer_${TAG}_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case: done' gpurun_out/sanitizer_${TAG}_$tool.log | tr '\n' ' ')"
done
