#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU): launch list of one bench cycle + full capture of the persistent kernel
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:ppo_persist -s 1 -c 1 \
    -o gpurun_out/r2_ppo_persist python tools/persist_once.py 256 > gpurun_out/r2_ncu_persist.log 2>&1
ncu -i gpurun_out/r2_ppo_persist.ncu-rep --page raw --csv > gpurun_out/r2_ppo_persist_raw.csv 2>/dev/null
tail -3 gpurun_out/r2_ncu_persist.log
