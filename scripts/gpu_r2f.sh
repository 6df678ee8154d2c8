#!/bin/bash
# round 2, call F: second-order sensitivity kernel for the initial-condition gradient (ros_ic_kernel)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "rosenbrock or stiff or robertson or small_batch or second_order" > $O/pytest_r2f.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_r2f.log
rm -f $O/stiff_r2f.jsonl
timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | grep chem6 | tee -a $O/stiff_r2f.jsonl
