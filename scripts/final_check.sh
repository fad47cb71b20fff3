#!/bin/bash
# end-of-round verification on one B200: GPU tests, smoke, every workload's bench line (e2e, roofline, cpu_baseline), the
# default bench line, and the time-only launch lists of the headline and the config-3 bench commands
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
for w in viterbi_k7_n1024_hard viterbi_c2 turbo_c3 ldpc_c4 link_c5; do
  timeout 600 python bench.py --workload $w --no-extras > gpurun_out/bench_r02b_$w.json 2> gpurun_out/bench_r02b_$w.err || echo "bench $w failed"
done
timeout 600 python bench.py > gpurun_out/bench_r02b_default.json 2> gpurun_out/bench_r02b_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_turbo_launches.csv python bench.py --workload turbo_c3 --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02b_")[1], "%.4g" % d["value"], d["unit"], "ms/step %.3f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"],
              "e2e %.4g" % d["e2e"]["value"], "cpu %.4g" % d.get("cpu_baseline", {}).get("value", 0), d.get("parity"), d["clocks"].get("reasons"))
    except Exception as e:
        print(f, "unreadable", e)
PY
