#!/bin/bash
# ncu captures of the kernels added / changed late in round 1 (one launch each, full set)
set -x
ncu --set full --clock-control none --import-source on -k regex:cn_bulk -s 1 -c 1 -o gpurun_out/prof_ldpc_bulk -f python scripts/profile_decoders.py ldpc > gpurun_out/ncu_ldpc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vn_kernel -s 1 -c 1 -o gpurun_out/prof_ldpc_vn2 -f python scripts/profile_decoders.py ldpc >> gpurun_out/ncu_ldpc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:demod_soft_separable -s 1 -c 1 -o gpurun_out/prof_demap2 -f python scripts/profile_decoders.py demap > gpurun_out/ncu_demap.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_link_tx -c 1 -o gpurun_out/prof_tx -f python scripts/profile_decoders.py tx > gpurun_out/ncu_tx.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ldpc_launches_final.csv python scripts/profile_decoders.py ldpc > /dev/null 2>&1
