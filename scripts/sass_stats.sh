#!/bin/bash
# Compile the LV-fishing-only library and print SASS statistics of the RK4 gradient kernel.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$ROOT/gpurun_out/sass          # scratch (git-ignored)
mkdir -p $W
cd $W
printf '#pragma once\n#include "%s/dynamicoed.jl_b200/csrc/generated/model_lv_fishing.cuh"\n#define DYNOED_FOR_EACH_MODEL(X) X(Model_lv_fishing)\n' "$ROOT" > models.cuh
if [ -z "$NOBUILD" ]; then nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v \
  -I$ROOT/include "-DDYNOED_MODELS_HEADER=\"$W/models.cuh\"" $EXTRA_DEFINES -o lv.so $ROOT/dynamicoed.jl_b200/csrc/dynoed_api.cu 2> ptxas.log; fi
grep -A1 "oed_eval_kernelI16Model_lv_fishingNS_6TabRK4ELb1ELb1" ptxas.log | grep -E "registers|spill" | head -4
cuobjdump -sass -fun '_ZN6dynoed15oed_eval_kernelI16Model_lv_fishingNS_6TabRK4ELb1ELb1EEEvNS_10EvalParamsE' lv.so > k.sass
python3 - <<'PY'
import re,collections
ins=[]
for l in open('k.sass'):
    m=re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
print('total instr',len(ins))
addr={a:i for i,(a,_) in enumerate(ins)}
def nreg(x):
    t=x.split()
    if t[0].startswith('@'): t=t[1:]
    ops=' '.join(t[1:])
    srcs=[o.strip() for o in ops.split(',')][1:]
    regs=set()
    for o in srcs:
        o=o.replace('-','').replace('|','').split('.')[0]
        if re.match(r'R\d+$',o): regs.add(o)
    return len(regs)
for i,(a,s) in enumerate(ins):
    m=re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)',s)
    if m and int(m.group(1),16)<a and int(m.group(1),16) in addr:
        j=addr[int(m.group(1),16)]
        body=ins[j:i+1]
        if len(body)<60: continue
        fp=[x for _,x in body if re.match(r'(@!?U?P\d+\s+)?D(FMA|MUL|ADD)',x)]
        c=collections.Counter(nreg(x) for x in fp)
        cyc=sum((3 if k>=3 else 2)*v for k,v in c.items())
        oth=collections.Counter(x.split()[1 if x.startswith('@') else 0].split('.')[0] for _,x in body if x not in fp)
        print(f'loop {j}-{i} len {len(body)} fp64 {len(fp)} by #distinct reg srcs {dict(sorted(c.items()))} est pipe cycles {cyc} other {len(body)-len(fp)} {dict(oth.most_common(8))}')
PY
