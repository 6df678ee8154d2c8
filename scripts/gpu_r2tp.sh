#!/bin/bash
# evidence pass after the time-parallel small-batch kernel: gpu_round.sh + the tp probes + ncu of tp_eval_kernel.
# scripts/gpu_r2tp.sh <tag> [timing variant]
TAG=${1:-r2tp}; O=gpurun_out; mkdir -p $O
bash scripts/gpu_round.sh $TAG
timeout 300 python scripts/tp_probe.py > $O/tp_probe_$TAG.txt 2>&1; echo "tp probe rc=$?"; tail -20 $O/tp_probe_$TAG.txt
timeout 200 python scripts/latency_probe.py > $O/latency_$TAG.txt 2>&1; echo "latency rc=$?"; cat $O/latency_$TAG.txt
[ -n "$2" ] && bash scripts/gpu_tp_phase.sh $TAG $2
timeout 200 python scripts/prof_dir.py > $O/plain_tp_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tp_eval -s 2 -c 1 -f -o $O/prof_tp_$TAG \
    python scripts/prof_dir.py > $O/ncu_tp_$TAG.log 2>&1
echo "tp capture rc=$?"; cat $O/plain_tp_$TAG.log
