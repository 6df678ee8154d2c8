#!/bin/bash
# round 2, call H: what would less checkpoint traffic buy?  A/B of the LV-only builds: default, DYNOED_TRAJ_WRAP
# (no checkpoint bytes to HBM at all: the bound for any recompute-instead-of-store scheme; results wrong),
# DYNOED_NO_L2_POLICY (plain stores/loads), and the L2 hot-fraction sweep on the default build.
O=gpurun_out; mkdir -p $O
V=dynamicoed.jl_b200/_variants
rm -f $O/ab_r2h.jsonl
for lib in base wrap nopol base wrap; do
  DYNOED_LIB=$V/libdynoed_$lib.so timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a $O/ab_r2h.jsonl
done
for f in 0 0.5 2.0; do
  DYNOED_L2_HOT_FRAC=$f DYNOED_LIB=$V/libdynoed_base.so timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a $O/ab_r2h.jsonl
done
