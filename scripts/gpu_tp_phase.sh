#!/bin/bash
# phase cycle counts of tp_eval_kernel from DYNOED_TP_TIMING variant builds.   scripts/gpu_tp_phase.sh <tag> <variant>...
O=gpurun_out; mkdir -p $O; TAG=${1:-a}; shift
for V in "$@"; do DYNOED_LIB=dynamicoed.jl_b200/_variants/libdynoed_$V.so timeout 200 python scripts/tp_phase_probe.py 2>&1 | tail -4; done | tee $O/tp_phase_$TAG.txt
