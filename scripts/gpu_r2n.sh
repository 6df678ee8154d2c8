#!/bin/bash
# round 2, call N: ncu --set full of ONE 1,048,576-design launch of the headline kernel (18 waves: CTAs in
# mixed phases, the closest a serialised ncu capture gets to the back-to-back steady state of bench.py)
O=gpurun_out; mkdir -p $O
timeout 200 python scripts/prof_kernel.py 1048576 > $O/plain_prof_r2n.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oed_eval -s 3 -c 1 -f -o $O/prof_r2n \
    python scripts/prof_kernel.py 1048576 > $O/ncu_full_r2n.log 2>&1
echo "full capture rc=$?"; cat $O/plain_prof_r2n.log
