#!/bin/bash
# One-box evidence pass: GPU tests, the default bench line of both arms, the ncu launch list of the bench command and
# one full capture of the plan kernel. Usage: tools/evidence.sh <tag>   (outputs under gpurun_out/<tag>_*)
T=${1:-ev}; O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${T}_pytest.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 --preroll 80 --no-cpu-baseline > $O/${T}_bench_n1_contact.json 2> $O/${T}_bench_n1_contact.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_agent_plan --launch-skip 24 -c 2 -f -o $O/${T}_full_plan \
  python tools/profile_steps.py 16 > $O/${T}_full_plan.log 2>&1
tail -3 $O/${T}_pytest.log; cat $O/${T}_bench_n1.json | cut -c1-1500
