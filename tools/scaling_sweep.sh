#!/bin/bash
# agent-replans/s at N in {16,64,256,512,1024} on the GPUs of this box. Usage: tools/scaling_sweep.sh <gpus> <out.jsonl>
G=${1:-1}; OUT=${2:-gpurun_out/sweep.jsonl}; : > "$OUT"
for W in circle circle_forest; do
  for N in 16 64 256 512 1024; do
    if [ "$G" = "1" ]; then
      python bench.py --workload $W --agents $N --steps 100 --warmup 10 --no-cpu-baseline >> "$OUT" 2>/dev/null
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 \
        bench.py --gpus $G --workload $W --agents $N --steps 100 --warmup 10 2>/dev/null | grep '^{' >> "$OUT"
    fi
  done
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(f"{d['config']['workload']:22s} gpus {d['n_gpus']} value {d['value']:12.0f} e2e {d['e2e']['value']:12.0f} ms/step {d['ms_per_step']:.3f} qp_fail {d['qp']['failed_last_step']}")
PY
