#!/bin/bash
# Round-2 evidence (one B200 under gpurun): launch list of the exact bench command + full captures of the two Viterbi kernels.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
EXP_REPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast_kernel_hard -s 2 -c 1 \
    -o gpurun_out/r02_hard_final -f python scripts/exp_hard.py > gpurun_out/ncu_hard_final.log 2>&1
EXP_REPS=2 EXP_MODE=soft EXP_NBITS=4096 timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast_kernel_soft -s 2 -c 1 \
    -o gpurun_out/r02_soft_c2_final -f python scripts/exp_hard.py > gpurun_out/ncu_soft_final.log 2>&1
tail -2 gpurun_out/ncu_hard_final.log gpurun_out/ncu_soft_final.log
