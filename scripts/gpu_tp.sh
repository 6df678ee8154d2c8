#!/bin/bash
# time-parallel small-batch kernel: probe (+ optional timing variants); TESTS=1 adds the small-batch tests.   scripts/gpu_tp.sh <tag> [variant...]
TAG=${1:-tp}; O=gpurun_out; mkdir -p $O; shift
timeout 300 python scripts/tp_probe.py > $O/tp_probe_$TAG.txt 2>&1; echo "probe rc=$?"; cat $O/tp_probe_$TAG.txt | tail -40
[ $# -gt 0 ] && bash scripts/gpu_tp_phase.sh $TAG "$@"
if [ -n "$TESTS" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "time_parallel or small_batch or hess or reference_ or rosenbrock or callback" > $O/pytest_tp_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_tp_$TAG.log; fi
