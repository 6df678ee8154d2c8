#!/bin/bash
# round 2, call G: full evidence pass on the committed build: GPU tests, bench (+ reference arm), launch list,
# ncu --set full of the headline kernel and of the chem6 second-order kernel, stiff probe, A/B probe.
O=gpurun_out; mkdir -p $O
bash scripts/gpu_round.sh r2g
timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | tee $O/stiff_r2g.jsonl | cut -c1-400
timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee $O/ab_r2g.jsonl
timeout 200 python scripts/prof_chem6.py > $O/plain_chem6_r2g.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ros_ -s 4 -c 2 -f -o $O/prof_chem6_r2g \
    python scripts/prof_chem6.py > $O/ncu_chem6_r2g.log 2>&1
echo "chem6 capture rc=$?"; cat $O/plain_chem6_r2g.log
