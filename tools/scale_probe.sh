set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_n1.err
for G in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2954$G bench.py --gpus $G --steps 20 --warmup 5 2>gpurun_out/r2u_n$G.err | grep "^{" > gpurun_out/r2u_bench_n$G.json
done
python - <<'PY'
import json
for g in (1,2,4):
    try:
        d=json.loads(open(f"gpurun_out/r2u_bench_n{g}.json").read().strip().splitlines()[-1])
        print(g, round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_step"].items()}, (d.get("sharded_check") or {}).get("single_engine_digest") == (d.get("sharded_check") or {}).get("sharded_digest"))
    except Exception as ex: print(g, "ERR", ex)
PY
