#!/bin/bash
# quick check of a build: GPU tests + the bench line.   scripts/gpu_quick.sh <tag>
TAG=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"
