#!/bin/bash
# ncu capture of the LDPC check pass (2nd launch) with source counters
ncu --set full --clock-control none --import-source on -k regex:cn_bulk -s 1 -c 1 -o gpurun_out/prof_ldpc_bulk -f python scripts/profile_decoders.py ldpc > gpurun_out/ncu_ldpc.log 2>&1
