"""Writes profiles/sass_tp_eval_lv_rk4.txt: static SASS opcode histogram of the time-parallel small-batch kernel
(tp_eval_kernel<Model_lv_fishing, ExplicitStepper<TabRK4>, SPILL=0>), from the LV-only build that
scripts/sass_stats.sh leaves in gpurun_out/sass/lv.so:
  cuobjdump -sass lv.so | awk '<the function>' > gpurun_out/sass/tp.sass   (see NOTES.md)"""
import collections, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ins = []
for l in open(os.path.join(ROOT, "gpurun_out", "sass", "tp.sass")):
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append(m.group(2).strip())
def op(x):
    t = x.split()
    return (t[1] if t[0].startswith('@') else t[0])
full = collections.Counter(op(x) for x in ins)
base = collections.Counter(op(x).split('.')[0] for x in ins)
fp = sum(base[k] for k in ('DFMA', 'DMUL', 'DADD'))
out = ["# SASS opcode histogram: tp_eval_kernel<Model_lv_fishing, ExplicitStepper<TabRK4>, SPILL=0> (sm_100a)",
       "# static instruction counts (cuobjdump -sass of an LV-only build of csrc/dynoed_api.cu); executed counts of one launch:",
       "# profiles/tp_r2tpf_summary.md (ncu source page)",
       f"# total {len(ins)} instructions; FP64 arithmetic {fp} = DFMA {base['DFMA']} / DMUL {base['DMUL']} / DADD {base['DADD']}",
       "# local memory: " + ", ".join(f"{k} x{v}" for k, v in sorted(full.items()) if re.match(r'LDL|STL', k)) + " (none if empty)",
       "# barriers / warp sync / sleep: " + ", ".join(f"{k} x{v}" for k, v in sorted(full.items())
                                                         if re.match(r'BAR|WARPSYNC|NANOSLEEP|SHFL|MEMBAR|ATOM|RED', k)),
       "", "## by base opcode"]
out += [f"{v:6d}  {k}" for k, v in base.most_common()]
out += ["", "## by full opcode (modifiers kept)"]
out += [f"{v:6d}  {k}" for k, v in full.most_common()]
open(os.path.join(ROOT, "profiles", "sass_tp_eval_lv_rk4.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:7]))
