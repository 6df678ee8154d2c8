#!/bin/bash
# round 2, call L: full GPU suite on the speculative-LU build + ncu of the chem6 kernels
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_r2l.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_r2l.log
timeout 200 python scripts/prof_chem6.py > $O/plain_chem6_r2l.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ros_ -s 4 -c 2 -f -o $O/prof_chem6_r2l \
    python scripts/prof_chem6.py > $O/ncu_chem6_r2l.log 2>&1
echo "chem6 capture rc=$?"; cat $O/plain_chem6_r2l.log
