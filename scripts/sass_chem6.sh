#!/bin/bash
# Compile a chem6-only library (no GPU needed) and print ptxas statistics + opcode mix of the
# Rosenbrock kernels: registers, stack frame, spills, FP64 share of the instruction stream.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$ROOT/gpurun_out/sass_chem6
mkdir -p $W; cd $W
printf '#pragma once\n#include "%s/dynamicoed.jl_b200/csrc/generated/model_chem6.cuh"\n#define DYNOED_FOR_EACH_MODEL(X) X(Model_chem6)\n' "$ROOT" > models.cuh
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v \
  -I$ROOT/include "-DDYNOED_MODELS_HEADER=\"$W/models.cuh\"" $EXTRA_DEFINES -o c6.so $ROOT/dynamicoed.jl_b200/csrc/dynoed_api.cu 2> ptxas.log
python3 - <<'PY'
import re, subprocess, collections
log = open('ptxas.log').read()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log)
for name, stack, ss, sl, regs in ents:
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if 'ros_eval_kernel' not in dem and 'ros_ic_kernel' not in dem: continue
    sass = subprocess.run(['cuobjdump', '-sass', '-fun', name, 'c6.so'], capture_output=True, text=True).stdout
    c = collections.Counter()
    for l in sass.splitlines():
        m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)', l)
        if m: c[m.group(1).split('.')[0]] += 1
    tot = sum(c.values()); fp = c['DFMA'] + c['DMUL'] + c['DADD']
    short = re.sub(r'dynoed::|void |\(dynoed::EvalParams\)', '', dem)
    print(f"{short}: regs {regs} stack {stack} spill st/ld {ss}/{sl} | instr {tot} fp64 {fp} ({100*fp/max(tot,1):.0f}%) DMUL {c['DMUL']} LDL {c['LDL']} STL {c['STL']} SEL {c['SEL']+c['FSEL']} BRA {c['BRA']}")
PY
