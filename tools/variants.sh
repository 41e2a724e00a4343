#!/bin/bash
# A/B of the engine's scheduling knobs on one box. Usage: tools/variants.sh [workload] [agents] -- VAR=.. VAR=.. / VAR=..
W=${1:-circle_forest}; N=${2:-1024}; shift 2
run() { echo "== $*"; env "$@" python bench.py --workload $W --agents $N --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print(f\"value {d['value']:.0f} e2e {d['e2e']['value']:.0f} ms/step {d['ms_per_step']:.4f} | lsc {k['k_lsc_build']:.4f} qp {k['k_qp_solve']:.4f} sfc {k['k_sfc_expand']:.4f} pred {k['k_predict']:.4f} | fail {d['qp']['failed_last_step']}\")"; }
for cfg in "$@"; do run ${cfg//,/ }; done
