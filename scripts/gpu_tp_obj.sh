#!/bin/bash
# objective-only calls on the time-parallel kernel: tests that touch small batches + the latency probe.   scripts/gpu_tp_obj.sh <tag>
TAG=${1:-obj}; O=gpurun_out; mkdir -p $O
timeout 100 python -m pytest tests -m gpu -x -q -k "time_parallel or ragged or small_batch or predictor or reference_lv" > $O/pytest_obj_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_obj_$TAG.log
timeout 40 python scripts/latency_probe.py > $O/latency_$TAG.txt 2>&1; echo "latency rc=$?"; cat $O/latency_$TAG.txt
