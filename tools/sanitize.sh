#!/bin/bash
# compute-sanitizer passes over the round-2 kernels. Usage: tools/sanitize.sh <out.txt>
OUT=${1:-gpurun_out/sanitizer.txt}; : > $OUT
run() { echo "== $*" >> $OUT; timeout 600 "$@" 2>&1 | grep -E "SUMMARY|ERROR SUMMARY|hazard|Invalid|Error|error:" | head -20 >> $OUT; }
S=/usr/local/cuda/bin/compute-sanitizer
run $S --tool memcheck python tools/goal_diag.py 3
run $S --tool racecheck python tools/goal_diag.py 2
run $S --tool synccheck python tools/goal_diag.py 2
run $S --tool memcheck python tools/gpu_diag.py --agents 1024 --steps 4 --every 4
run $S --tool racecheck python tools/gpu_diag.py --agents 300 --steps 5 --every 5
run $S --tool synccheck python tools/gpu_diag.py --agents 300 --steps 4 --every 4
run $S --tool memcheck python -m pytest tests/test_gpu_edges.py tests/test_gpu_slack.py -x -q
cat $OUT
