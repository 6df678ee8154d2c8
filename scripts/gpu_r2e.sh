#!/bin/bash
# round 2, call E: register-resident pivoted LU (select swaps + warp-uniform skip) in the Rosenbrock kernels
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "rosenbrock or stiff or robertson or small_batch" > $O/pytest_r2e.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2e.log
rm -f $O/stiff_r2e.jsonl
timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | tee -a $O/stiff_r2e.jsonl
