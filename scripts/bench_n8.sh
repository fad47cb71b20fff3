#!/bin/bash
# every BASELINE config at N GPUs of one box (default 8): one JSON line per workload into gpurun_out/bench_r02c_n<N>_<workload>.json
N=${1:-8}
mkdir -p gpurun_out
port=29510
for w in viterbi_k7_n1024_hard turbo_c3 ldpc_c4 link_c5 viterbi_c2; do
  port=$((port + 1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --workload $w --no-extras --steps 10 --warmup 3 > gpurun_out/bench_r02c_n${N}_$w.json 2> gpurun_out/bench_r02c_n${N}_$w.err || echo "bench $w failed"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02c_n${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("_n${N}_")[1], "n_gpus", d["n_gpus"], "%.4g" % d["value"], d["unit"], "ms/step %.3f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], d.get("result", {}).get("ber_per_point", ""))
    except Exception as e:
        print(f, "unreadable", e)
PY
