#!/bin/bash
# one gpurun call: turbo timing for the kernel / layout switches, the whole GPU test suite, and full ncu captures of the two
# MAP launches of one turbo iteration
mkdir -p gpurun_out
python scripts/exp_map.py > gpurun_out/exp_map.log 2>&1
cat gpurun_out/exp_map.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_lin2 -c 2 -o gpurun_out/r02_map_lin2 -f \
    python scripts/profile_decoders.py turbo > gpurun_out/ncu_map.log 2>&1
tail -3 gpurun_out/ncu_map.log
