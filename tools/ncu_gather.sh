#!/bin/bash
# one `ncu --set full` capture of the gather-mode forward (critic first layer) with per-source-line samples
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --clock-control none --set full --import-source on -k regex:conv_tc_kmajor --launch-skip 2 --launch-count 1 -f -o /tmp/r02_gather \
    python tools/gather_time.py > gpurun_out/r02_gather.log 2>&1
ncu -i /tmp/r02_gather.ncu-rep --page raw --csv > gpurun_out/r02_gather_raw.csv 2>/dev/null
ncu -i /tmp/r02_gather.ncu-rep --page source --csv 2>/dev/null | head -c 8000000 > gpurun_out/r02_gather_source.csv
ls -la gpurun_out/r02_gather*
