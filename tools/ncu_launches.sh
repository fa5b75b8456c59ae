#!/bin/bash
# Launch lists only (gpu__time_duration.sum) of one eager update_model for BASELINE configs[2] (bench default) and configs[1],
# plus one `ncu --set full` capture of the small-channel direct conv kernels.  Outputs under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --ncu --warmup 1 --steps 1 > gpurun_out/r02_ncu_c3.json 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --ncu --config 2 --warmup 1 --steps 1 > gpurun_out/r02_ncu_c2.json 2>&1
$NCU --set full --import-source on -k "regex:small_" --launch-skip 9 --launch-count 9 -f -o /tmp/r02_small python bench.py --ncu --warmup 1 --steps 1 > gpurun_out/r02_small.log 2>&1
ncu -i /tmp/r02_small.ncu-rep --page raw --csv > gpurun_out/r02_small_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv | tail -5
