#!/bin/bash
# round 2, call K: speculative no-exchange LU in the Rosenbrock kernels: parity subset + stiff probe
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "rosenbrock or stiff or robertson or small_batch or second_order" > $O/pytest_r2k.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2k.log
timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | tee $O/stiff_r2k.jsonl | cut -c1-330
