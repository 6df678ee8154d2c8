#!/bin/bash
# GPU suite + stiff probe (tag = $1)
TAG=${1:-r2r}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_${TAG}.log
timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | tee $O/stiff_${TAG}.jsonl | cut -c1-60
