#!/bin/bash
# end-of-round verification on one B200: GPU tests, smoke, bench, launch list of the bench command, soft-kernel profile
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 3 --warmup 3 --no-extras > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast_kernel_soft -s 1 -c 1 -o gpurun_out/prof_viterbi_soft2 -f python scripts/profile_soft.py > gpurun_out/ncu_soft.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final4.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"])
for k,v in d["extras"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms","roofline_frac","ber","seconds","tx_ms_per_batch","error")})
PY
