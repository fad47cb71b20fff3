mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_links.py tests/test_decoders_gpu.py -m gpu -x -q 2>&1 | tail -4
python scripts/exp_link.py 2>&1 | tail -5
for w in turbo_c3 ldpc_c4 link_c5; do
  timeout 600 python bench.py --workload $w --no-extras > gpurun_out/bench_r02c_$w.json 2> gpurun_out/bench_r02c_$w.err || echo "bench $w failed"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02c_")[1], "%.4g" % d["value"], d["unit"], "ms/step %.3f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
