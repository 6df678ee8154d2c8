"""A/B timing of the LV-fishing RK4 gradient kernel for the library named by DYNOED_LIB (default: the
in-tree build): back-to-back (overlapped) and isolated launches of 65,536 designs on rotating buffers,
plus clocks/power under load.  Prints one JSON line."""
import json, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

n, NB = 65536, 6
prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
ctx = prob.context()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
bufs = []
for b in range(NB):
    rng = np.random.Generator(np.random.Philox(20262 + b))
    h = rng.random((n, 71)); h[:, -1] = 1e-3 + (1 - 1e-3) * h[:, -1]
    d = torch.from_numpy(h).cuda()
    bufs.append((d, torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty_like(d),
                 torch.empty((n, 3), dtype=torch.float64, device="cuda")))
def run(k, reps, grad=True):
    for i in range(reps):
        d, c, g, f = bufs[(k + i) % NB]
        ctx.eval_batch(d, c, g if grad else None, f, async_=True, flags=2)
run(0, 10); torch.cuda.synchronize()
rows = []
def sample():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    sample.p = p
    for line in p.stdout:
        rows.append(line.strip())
th = threading.Thread(target=sample, daemon=True); th.start(); time.sleep(0.2)
out = {"lib": os.path.basename(runtime.LIB_PATH), "hot_frac": os.environ.get("DYNOED_L2_HOT_FRAC")}
for grad in (True, False):
    reps = 2000 if grad else 4000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(0, reps, grad); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ks = ctx.kernel_stats(grad)
    fl = ks["flops_per_design_grad" if grad else "flops_per_design_eval"]
    out["grad" if grad else "eval"] = {"ms": ms, "Mdesigns_s": n / ms / 1e3, "tflops": fl * n / ms / 1e9,
                                       "flops_per_design": fl, "regs": ks["regs_per_thread"]}
sample.p.terminate()
sm = [float(r.split(",")[0]) for r in rows if "," in r]; pw = [float(r.split(",")[1]) for r in rows if "," in r]
load = [s for s, p in zip(sm, pw) if p > 0.6 * max(pw)]
out["sm_mhz_under_load"] = float(np.median(load)) if load else None
out["power_w_max"] = max(pw) if pw else None
ctx.set_profiling(True); run(0, 30); torch.cuda.synchronize()
out["isolated_ms"] = ctx.kernel_time()[0]; ctx.set_profiling(False)
out["crit0"] = float(bufs[0][1][0])
# order-independent fingerprints of buffer set 0 (bit-identical between builds that do the same arithmetic)
out["crit_xor"] = int(torch.bitwise_xor(bufs[0][1].view(torch.int64)[::2], bufs[0][1].view(torch.int64)[1::2]).sum())
out["grad_sum"] = float(bufs[0][2].sum())
out["fp64_peak"] = runtime.fp64_peak()[0]
print(json.dumps(out))
