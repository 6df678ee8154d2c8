"""HBM-streaming fixed-controls kernel: achieved bandwidth at 1M designs."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
base = np.random.default_rng(0).random(71)
basis = torch.from_numpy(ctx.fim_basis(base)).cuda()
n = 1 << 20
bufs = [torch.rand((n, 71), dtype=torch.float64, device="cuda") for _ in range(2)]
c = torch.empty(n, dtype=torch.float64, device="cuda"); g = [torch.empty_like(bufs[0]) for _ in range(2)]
for k in range(4): ctx.eval_fixed_controls(basis, bufs[k % 2], c, g[k % 2], None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(20): ctx.eval_fixed_controls(basis, bufs[k % 2], c, g[k % 2], None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
byt = n * 8 * (71 * 2 + 1)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6532.2
print(f"fixed_controls_kernel n={n}: {ms*1e3:.1f} us, {n/ms/1e6:.1f} Gdesigns/s, {byt/ms/1e6:.0f} GB/s = {byt/ms/1e6/peak:.2f} of measured HBM copy bandwidth ({peak} GB/s)")
