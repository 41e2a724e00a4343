#!/bin/bash
# A/B of the split step (wide blocks for the most expensive agents). Usage: tools/split_sweep.sh "<threads>:<k> ..." [bench flags]
CFGS=$1; shift
for c in $CFGS; do
  T=${c%%:*}; K=${c##*:}
  LSCGPU_SPLIT_K=$K LSCGPU_SPLIT_THREADS=$T timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('threads $T k $K value %.0f e2e %.0f ms/step %.4f plan %.4f same_traj %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], k['k_agent_plan'], d['e2e']['same_trajectories_as_resident_pass']))"
done
