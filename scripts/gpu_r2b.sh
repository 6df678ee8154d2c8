#!/bin/bash
# round 2, call B: tests of the guard / opt-in overlap / selection kernels, L2-policy sweep, bench.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_r2b.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2b.log
V=dynamicoed.jl_b200/_variants
rm -f $O/ab_r2b.jsonl
DYNOED_LIB=$V/libdynoed_nopol.so timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a $O/ab_r2b.jsonl
for f in 0 0.25 0.5 0.75 1.0 1.5; do
  DYNOED_L2_HOT_FRAC=$f timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a $O/ab_r2b.jsonl
done
timeout 900 python bench.py --steps 200 --warmup 5 > $O/bench_r2b.log 2> $O/bench_r2b.err; echo "bench rc=$?"; cat $O/bench_r2b.log; tail -5 $O/bench_r2b.err
