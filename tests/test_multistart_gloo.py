"""N>1 host logic (sharding, (min, index) gather, winner broadcast) with world_size=2 over gloo
on CPU.  The evaluator is injected: here the CPU oracle stands in for the GPU library."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynamicoed.jl_b200 import models
from dynamicoed.jl_b200.multistart import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 65536, 1048576 + 3):
        for w in (1, 2, 4, 8):
            parts = [shard_bounds(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= \
                max(1, -(-n // w))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicoed.jl_b200.multistart import sweep
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle, FISHER_D
    o = Oracle(models.lotka_volterra_fishing(), FISHER_D, "rk4", 2)
    c = CppOracle("lv_fishing", o)
    rng = np.random.Generator(np.random.Philox(99))
    n = 301                                     # ragged: 151 + 150
    des = rng.random((n, 71))
    lo, hi = shard_bounds(n, rank, world)

    def evaluate(d):
        cr, g, _ = c.evaluate(d.numpy(), grad=True, threads=2)
        k = int(np.argmin(cr))
        return torch.from_numpy(cr), torch.from_numpy(g), (k, float(cr[k]))
    out = sweep(evaluate, torch.from_numpy(des[lo:hi]), lo, gather_crit=True)
    cr_all, g_all, _ = c.evaluate(des, grad=True, threads=2)
    k = int(np.argmin(cr_all))
    ok = (out["index"] == k and out["value"] == cr_all[k] and
          np.array_equal(out["design"].numpy(), des[k]) and
          np.array_equal(out["grad"].numpy(), g_all[k]) and
          np.array_equal(out["crit_all"].numpy(), cr_all) and out["owner"] == (0 if k < hi - lo or rank == 0 and k < hi else out["owner"]))
    # the device-side protocol (one all_gather of packed local winners, no host round trips): the
    # shard's per-batch records, here two "batches" per rank, packed and reduced with the same code
    from dynamicoed.jl_b200.multistart import gather_winner, pack_local_winner
    cr_loc, g_loc = torch.from_numpy(cr_all[lo:hi].copy()), torch.from_numpy(g_all[lo:hi].copy())
    half = (hi - lo) // 2
    vals = torch.stack([cr_loc[:half].min(), cr_loc[half:].min()])
    idxs = torch.stack([cr_loc[:half].argmin(), cr_loc[half:].argmin() + half])
    win = gather_winner(pack_local_winner(vals, idxs, lo, torch.from_numpy(des[lo:hi]), g_loc))
    ok = ok and (int(win[1]) == k and float(win[0]) == cr_all[k] and int(win[-1]) == out["owner"] and
                 np.array_equal(win[2:73].numpy(), des[k]) and np.array_equal(win[73:144].numpy(), g_all[k]))
    # a rank without any valid candidate (all NaN) never wins; if nobody has one: index -1, owner -1
    nanp = pack_local_winner(torch.tensor([float("nan")], dtype=torch.float64), torch.tensor([5]), lo,
                             torch.from_numpy(des[lo:hi]), g_loc)
    win2 = gather_winner(nanp if rank == 0 else pack_local_winner(vals, idxs, lo, torch.from_numpy(des[lo:hi]), g_loc))
    k1 = lo_hi_best = None
    lo1, hi1 = shard_bounds(n, 1, world)
    k1 = lo1 + int(np.argmin(cr_all[lo1:hi1]))
    ok = ok and int(win2[1]) == k1 and int(win2[-1]) == 1
    win3 = gather_winner(nanp)
    ok = ok and int(win3[1]) == -1 and int(win3[-1]) == -1 and float(win3[0]) == float("inf")
    q.put((rank, bool(ok), out["index"], k))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_sweep_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res


def test_pack_and_gather_winner_single_process_edges():
    from dynamicoed.jl_b200.multistart import gather_winner, pack_local_winner
    rng = np.random.default_rng(3)
    d = torch.from_numpy(rng.random((10, 5))); g = torch.from_numpy(rng.random((10, 5)))
    # NaN never wins, ties go to the smallest index, indices < 0 mean "no candidate in that batch"
    vals = torch.tensor([0.5, float("nan"), -1.0, -1.0, -7.0])
    idx = torch.tensor([3, 4, 7, 2, -1])
    p = pack_local_winner(vals, idx, 100, d, g)
    assert float(p[0]) == -1.0 and int(p[1]) == 102
    assert torch.equal(p[2:7], d[2]) and torch.equal(p[7:], g[2])
    w = gather_winner(p)
    assert float(w[0]) == -1.0 and int(w[1]) == 102 and int(w[-1]) == 0 and torch.equal(w[:-1], p)
    # -inf is a legitimate value
    p = pack_local_winner(torch.tensor([float("-inf"), 1.0]), torch.tensor([9, 0]), 0, d, g)
    assert float(p[0]) == float("-inf") and int(p[1]) == 9 and torch.equal(p[2:7], d[9])
    # no candidate at all / empty shard
    for v, i, dd, gg in ((torch.tensor([float("nan")]), torch.tensor([5]), d, g),
                         (torch.tensor([float("inf")]), torch.tensor([-1]), d[:0], g[:0])):
        p = pack_local_winner(v, i, 100, dd, gg)
        assert float(p[0]) == float("inf") and int(p[1]) == -1 and float(p[2:].abs().sum()) == 0.0
        w = gather_winner(p)
        assert int(w[1]) == -1 and int(w[-1]) == -1
