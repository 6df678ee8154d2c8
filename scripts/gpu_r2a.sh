#!/bin/bash
# round 2, call A: parity of the reformulated kernel, A/B against the round-1 kernel and the
# no-checkpoint-traffic bound, bench line.
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_r2a.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2a.log
V=dynamicoed.jl_b200/_variants
for lib in "" $V/libdynoed_old.so $V/libdynoed_wrap.so ""; do
  DYNOED_LIB=$lib timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a $O/ab_r2a.jsonl
done
timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench_r2a.log 2> $O/bench_r2a.err; echo "bench rc=$?"; cat $O/bench_r2a.log
