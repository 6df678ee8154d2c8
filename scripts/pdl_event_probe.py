"""Does a cudaEventRecord between two launches prevent programmatic overlap?  (design probe)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n, nv, NB = 65536, 71, 4
ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
S = torch.cuda.current_stream(); ctx.set_stream(S.cuda_stream)
T = torch.cuda.Stream()
ins = [torch.rand((n, nv), dtype=torch.float64, device="cuda") for _ in range(NB)]
outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty((n, nv), dtype=torch.float64, device="cuda"),
         torch.empty((n, 3), dtype=torch.float64, device="cuda")) for _ in range(NB)]
rec = torch.empty(2, dtype=torch.int64, device="cuda")
def run(mode, K=200):
    evs = [torch.cuda.Event() for _ in range(K)]
    for _ in range(5): ctx.eval_batch(ins[0], *outs[0], async_=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(S)
    for k in range(K):
        ctx.eval_batch(ins[k % NB], *outs[k % NB], async_=True)
        if mode == "event":
            evs[k].record(S)
        elif mode == "event+side":
            evs[k].record(S); T.wait_event(evs[k])
            with torch.cuda.stream(T): rec.add_(1)
        elif mode == "copy":
            ctx.argmin_record_async(rec)
    e1.record(S); torch.cuda.synchronize()
    print(f"{mode:12s} {e0.elapsed_time(e1)/K*1e3:8.1f} us/step", flush=True)
for m in ("plain", "event", "event+side", "copy", "plain"):
    run(m)
