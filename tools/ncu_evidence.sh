#!/bin/bash
# Round-2 profiling evidence (run on the GPU box through gpurun; outputs under gpurun_out/):
#   1. launch lists (gpu__time_duration.sum) of one eager update_model for BASELINE configs[2] (the bench default) and configs[1]
#   2. `ncu --set full` captures of the dominant kernels of the 14-class step
# Numbers printed by a process running under ncu are never bench values.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --ncu --warmup 1 --steps 1 > gpurun_out/r02_ncu_c3.json 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --ncu --config 2 --warmup 1 --steps 1 > gpurun_out/r02_ncu_c2.json 2>&1
# full captures: kernel-name filtered, a few instances each, from the second (warm) step of the 14-class run
# (the .ncu-rep files stay on the box: gpurun brings back at most 64 MiB; what travels is the raw-metric CSV of every captured
# launch and, for the first launch, the per-source-line page)
cap() {  # name regex skip count
  $NCU --set full --import-source on -k "regex:$2" --launch-skip "$3" --launch-count "$4" -f -o "/tmp/r02_$1" \
      python bench.py --ncu --warmup 1 --steps 1 > "gpurun_out/r02_$1.log" 2>&1
  ncu -i "/tmp/r02_$1.ncu-rep" --page raw --csv > "gpurun_out/r02_$1_raw.csv" 2>/dev/null
  ncu -i "/tmp/r02_$1.ncu-rep" --page source --csv --kernel-id :::1 2>/dev/null | head -c 3000000 > "gpurun_out/r02_$1_source.csv"
}
cap kmajor    'conv_tc_kmajor'   370 12
cap wgrad     'conv_tc_wgrad'    90 4
cap instnorm  'instnorm_.*_sm'   140 8
cap mrugate   'mru_gate_'        26 4
cap simt      'igemm_simt'       60 4
# the dominant layer class alone (roofline.traffic): DRAM bytes of one launch against its algorithmic bytes
for what in fwd dgrad wgrad; do
  $NCU --set full -k "regex:conv_tc_" --launch-skip 2 --launch-count 1 -f -o "/tmp/r02_conv3x3_$what" python tools/one_conv.py $what > "gpurun_out/r02_conv3x3_$what.log" 2>&1
  ncu -i "/tmp/r02_conv3x3_$what.ncu-rep" --page raw --csv > "gpurun_out/r02_conv3x3_${what}_raw.csv" 2>/dev/null
done
ls -la gpurun_out/ /tmp/*.ncu-rep; du -sh gpurun_out
