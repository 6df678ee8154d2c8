#!/bin/bash
# A/B of LV-only measurement builds: scripts/gpu_ab.sh <tag> <variant> [<variant> ...]   (each variant run once, in order)
O=gpurun_out; mkdir -p $O
TAG=$1; shift
V=dynamicoed.jl_b200/_variants
rm -f $O/ab_$TAG.jsonl
for lib in "$@"; do
  DYNOED_LIB=$V/libdynoed_$lib.so timeout 300 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a $O/ab_$TAG.jsonl
done
