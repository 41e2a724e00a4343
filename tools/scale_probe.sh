#!/bin/bash
# bench.py at 1 .. G GPUs of this box (driver-style flags). Usage: tools/scale_probe.sh <max gpus> <tag> [extra bench flags]
G=${1:-2}; TAG=${2:-probe}; shift 2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_n1.err
for N in 2 4 8; do
  [ $N -le $G ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>gpurun_out/${TAG}_n$N.err | grep "^{" > gpurun_out/${TAG}_bench_n$N.json
done
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
base = None
for g in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/{tag}_bench_n{g}.json").read().strip().splitlines()[-1])
    except Exception as ex:
        continue
    base = base or d["value"]
    sc = d.get("sharded_check") or {}
    print(g, round(d["value"]), "eff %.2f" % (d["value"] / base / g), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]),
          {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()}, "replay-identical", sc.get("single_engine_digest") == sc.get("sharded_digest"))
PY
