"""Writes profiles/sass_oed_eval_lv_rk4_grad.txt: static SASS opcode histogram of the headline kernel, from
the LV-only build that scripts/sass_stats.sh leaves in gpurun_out/sass/k.sass (run that first)."""
import collections, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ins = []
for l in open(os.path.join(ROOT, "gpurun_out", "sass", "k.sass")):
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append(m.group(2).strip())
def op(x):
    t = x.split()
    return (t[1] if t[0].startswith('@') else t[0])
full = collections.Counter(op(x) for x in ins)
base = collections.Counter(op(x).split('.')[0] for x in ins)
fp = sum(base[k] for k in ('DFMA', 'DMUL', 'DADD'))
out = ["# SASS opcode histogram: oed_eval_kernel<Model_lv_fishing, TabRK4, GRAD=1, UCOEF=1> (sm_100a)",
       "# static instruction counts of the compiled kernel (cuobjdump -sass of an LV-only build of csrc/dynoed_api.cu,",
       "# scripts/sass_stats.sh + scripts/sass_histogram.py); executed counts are in the r2*_summary.md files (ncu source page)",
       f"# total {len(ins)} instructions; FP64 arithmetic {fp} = DFMA {base['DFMA']} / DMUL {base['DMUL']} / DADD {base['DADD']}"
       f" -> DMUL share of FP64 instructions {100*base['DMUL']/fp:.1f} %",
       "# TMA / mbarrier / PDL evidence: " + ", ".join(f"{k} x{v}" for k, v in sorted(full.items())
                                                        if re.match(r'UBLKCP|SYNCS|UTMA|PREEXIT|CCTL|ACQBULK|FENCE|ELECT', k)),
       "", "## by base opcode"]
out += [f"{v:6d}  {k}" for k, v in base.most_common()]
out += ["", "## by full opcode (modifiers kept)"]
out += [f"{v:6d}  {k}" for k, v in full.most_common()]
open(os.path.join(ROOT, "profiles", "sass_oed_eval_lv_rk4_grad.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:6]))
