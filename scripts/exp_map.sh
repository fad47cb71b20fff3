#!/bin/bash
# one gpurun call: turbo timing, the whole GPU test suite, and the default bench line (all workloads as extras)
mkdir -p gpurun_out
python scripts/exp_map.py > gpurun_out/exp_map.log 2>&1
cat gpurun_out/exp_map.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err; head -c 1500 gpurun_out/bench_default.json
