#!/bin/bash
# evidence pass on a build (tag = $1): GPU tests, bench (+ reference arm), launch list, ncu --set full of
# the headline kernel (isolated 65,536-design launch and one 1,048,576-design launch) and of the chem6 kernels.
# The CSV pages of the two extra captures are exported on the box and their reports deleted (64 MiB return limit).
TAG=${1:-r2s}
O=gpurun_out; mkdir -p $O
bash scripts/gpu_round.sh ${TAG}
exp() { ncu -i $O/prof_$1.ncu-rep --page raw --csv > $O/prof_$1_raw.csv 2>/dev/null; ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1_source.csv 2>/dev/null; rm -f $O/prof_$1.ncu-rep; }
timeout 200 python scripts/prof_kernel.py 1048576 > $O/plain_prof_${TAG}1m.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oed_eval -s 3 -c 1 -f -o $O/prof_${TAG}1m \
    python scripts/prof_kernel.py 1048576 > $O/ncu_full_${TAG}1m.log 2>&1
echo "1M capture rc=$?"; cat $O/plain_prof_${TAG}1m.log; exp ${TAG}1m
timeout 200 python scripts/prof_chem6.py > $O/plain_chem6_${TAG}.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ros_ -s 4 -c 2 -f -o $O/prof_chem6_${TAG} \
    python scripts/prof_chem6.py > $O/ncu_chem6_${TAG}.log 2>&1
echo "chem6 capture rc=$?"; cat $O/plain_chem6_${TAG}.log; exp chem6_${TAG}
du -sh $O
