#!/bin/bash
# evidence pass after the time-parallel small-batch kernel: gpu_round.sh + the tp probes.   scripts/gpu_r2tp.sh <tag>
TAG=${1:-r2tp}; O=gpurun_out; mkdir -p $O
bash scripts/gpu_round.sh $TAG
timeout 300 python scripts/tp_probe.py > $O/tp_probe_$TAG.txt 2>&1; echo "tp probe rc=$?"; tail -20 $O/tp_probe_$TAG.txt
timeout 200 python scripts/latency_probe.py > $O/latency_$TAG.txt 2>&1; echo "latency rc=$?"; cat $O/latency_$TAG.txt
