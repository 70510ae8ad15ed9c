#!/bin/bash
# Per-launch duration of the fused training kernel (ncu gpu__time_duration, C2 shape, 50k rows), for A/B comparisons of
# kernel variants on one box: tools/duo_time.sh [repeats]
for i in $(seq 1 ${1:-2}); do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused_ --csv python tools/trace_fused.py 2>/dev/null \
    | grep -E "fused_(duo|mlp)" | awk -F'","' '{gsub(/"/,"",$NF); printf "%s ", $NF} END {print "us"}'
done
