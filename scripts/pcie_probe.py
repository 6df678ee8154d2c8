"""Host<->device copy ceiling of the e2e step on N GPUs of one box, one process per GPU (torchrun):
every rank moves the bytes of ONE 65,536-design step (37.2 MB of designs host->device, 39.3 MB of
objective + gradient + F device->host) from/to pinned host memory, both directions at once, no
kernel.  Reports per-rank and aggregate GB/s for
  plain   : pinned buffers allocated wherever the process happens to run
  numa    : the process is first bound to the CPUs of the GPU's NUMA node (sysfs), then the pinned
            buffers are allocated (first touch => node-local) and the copies are issued from there
so that the e2e scaling of bench.py (21 % efficiency at 8 GPUs) can be attributed to the
box (plain copies do not scale either) or to the host path (they do).

  python scripts/pcie_probe.py                                   # 1 GPU
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_probe.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamicoed.jl_b200.runtime import gpu_numa_cpus  # noqa: E402

N, NV, NF = 65536, 71, 3
REPS = 40


def run(dev, label):
    hp = torch.empty((N, NV), dtype=torch.float64).pin_memory(); hp.fill_(0.5)
    hc = torch.empty(N, dtype=torch.float64).pin_memory()
    hg = torch.empty((N, NV), dtype=torch.float64).pin_memory()
    hf = torch.empty((N, NF), dtype=torch.float64).pin_memory()
    dd = torch.empty((N, NV), dtype=torch.float64, device=dev)
    dc = torch.zeros(N, dtype=torch.float64, device=dev)
    dg = torch.zeros((N, NV), dtype=torch.float64, device=dev)
    df = torch.zeros((N, NF), dtype=torch.float64, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def once(h2d=True, d2h=True):
        if h2d:
            with torch.cuda.stream(s_in):
                dd.copy_(hp, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                hc.copy_(dc, non_blocking=True); hg.copy_(dg, non_blocking=True); hf.copy_(df, non_blocking=True)
    out = {}
    for mode, kw in (("both", {}), ("h2d_only", dict(d2h=False)), ("d2h_only", dict(h2d=False))):
        for _ in range(3):
            once(**kw)
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(REPS):
            once(**kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = (N * NV * 8 if kw.get("h2d", True) else 0) + (N * (1 + NV + NF) * 8 if kw.get("d2h", True) else 0)
        out[mode] = nbytes * REPS / dt / 1e9
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    res = {"plain": run(dev, "plain")}
    node, cpus = gpu_numa_cpus(local)
    if cpus:
        os.sched_setaffinity(0, cpus)
    res["numa"] = run(dev, "numa")
    res["numa_node"], res["n_cpus_bound"] = node, len(cpus or [])
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, res)
    else:
        rows = [res]
    if rank == 0:
        agg = {m: {k: sum(r[m][k] for r in rows) for k in ("both", "h2d_only", "d2h_only")} for m in ("plain", "numa")}
        print(json.dumps({"n_gpus": world, "host_cpus": os.cpu_count(), "bytes_per_step": N * NV * 8 + N * (1 + NV + NF) * 8,
                          "aggregate_gbs": agg, "designs_per_s_ceiling": {m: agg[m]["both"] * 1e9 / ((2 * NV + 1 + NF) * 8) for m in agg},
                          "per_rank": rows}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
