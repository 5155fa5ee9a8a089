#!/bin/bash
# Produces the ncu evidence kept under profiles/ (run under gpurun on one B200):
#   $1_launches.csv              per-launch durations of one config-2 batch (resident path, 2 iterations)
#   $1_full.ncu-rep / _raw.csv   `ncu --set full` of one launch of every kernel of the step (7 launches)
#   $1_<kernel>_src.csv          SASS-level source pages (tools/ncu_lines.py joins them with the line table)
tag=${1:-gpurun_out/prof}
GQ_OPTIONS=overlap_classify=0 GQ_PROFILE_ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file ${tag}_launches.csv python tools/profile_run.py > ${tag}_launches.log 2>&1
GQ_OPTIONS=overlap_classify=0 GQ_PROFILE_ITERS=2 ncu --set full --clock-control none --import-source on \
    -k regex:"revcomp_kernel|seed_kernel|verify_kernel|text_kernel|search_kernel|classify_kernel|coverage_kernel" -s 7 -c 7 \
    -o ${tag}_full -f python tools/profile_run.py > ${tag}_full.log 2>&1
ncu -i ${tag}_full.ncu-rep --page raw --csv > ${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py ${tag}_raw.csv
for k in seed_kernel text_kernel classify_kernel coverage_kernel verify_kernel revcomp_kernel; do
  ncu -i ${tag}_full.ncu-rep --page source --csv --kernel-name regex:$k > ${tag}_${k}_src.csv 2>/dev/null
done
