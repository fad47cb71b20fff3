#!/bin/bash
# soft-decision K=7 kernel: occupancy variants (scripts/build_variant.sh ... -DCPB_VITERBI_TBB_SOFT / CPB_TASK_CAP / CPB_SOFT_MIN_CTAS /
# CPB_SOFT_MAX_CARVEOUT) at config 2 (65,536 x 4096) and at 65,536 x 1024
for nb in 4096 1024; do
  EXP_MODE=soft EXP_NBITS=$nb EXP_REPS=5 python scripts/exp_hard.py
  for v in build/variants/libcommpy_b200_s*.so; do COMMPY_B200_LIB=$PWD/$v EXP_MODE=soft EXP_NBITS=$nb EXP_REPS=5 python scripts/exp_hard.py; done
done
