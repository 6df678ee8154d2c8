#!/bin/bash
# round 2, call R: 1/h step table in the Rosenbrock kernels: full GPU suite + stiff probe
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2r.log
timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | tee $O/stiff_r2r.jsonl | cut -c1-60
