#!/bin/bash
# last GPU call of the round: test suite, smoke, memcheck over the small-shape driver (new MAP / TX / transpose kernels included),
# full ncu capture of the final MAP kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/sanitize_memcheck_r02b.log 2>&1
echo "== memcheck: exit $?"; grep -E "ERROR SUMMARY|sanitize driver ok|Error:" gpurun_out/sanitize_memcheck_r02b.log | sort | uniq -c | head
timeout 100 ncu --set full --clock-control none --import-source on -k regex:map_lin2 -c 2 -o gpurun_out/r02_map_lin2_final -f \
    python scripts/profile_decoders.py turbo > gpurun_out/ncu_map_final.log 2>&1
tail -2 gpurun_out/ncu_map_final.log
