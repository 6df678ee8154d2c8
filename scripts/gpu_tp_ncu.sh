#!/bin/bash
# ncu --set full of ONE tp_eval_kernel launch (one LV design); the CSV pages are exported on the box and the
# report deleted (gpurun returns at most 64 MiB).   scripts/gpu_tp_ncu.sh <tag>
TAG=${1:-tp}; O=gpurun_out; mkdir -p $O
timeout 60 python scripts/prof_dir.py > $O/plain_tp_$TAG.log 2>&1 &&
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tp_eval -s 2 -c 1 -f -o $O/prof_tp_$TAG \
    python scripts/prof_dir.py > $O/ncu_tp_$TAG.log 2>&1
echo "tp capture rc=$?"; cat $O/plain_tp_$TAG.log
ncu -i $O/prof_tp_$TAG.ncu-rep --page raw --csv > $O/prof_tp_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_tp_$TAG.ncu-rep --page source --csv > $O/prof_tp_${TAG}_source.csv 2>/dev/null
rm -f $O/prof_tp_$TAG.ncu-rep; du -sh $O
