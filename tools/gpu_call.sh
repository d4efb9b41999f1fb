# scratch driver for one gpurun call (edited per call)
mkdir -p gpurun_out
python tools/sanitize_case_r02b.py > gpurun_out/s1_plain.log 2>&1; tail -n 2 gpurun_out/s1_plain.log
bash tools/run_sanitizer.sh r02b tools/sanitize_case_r02b.py 110
