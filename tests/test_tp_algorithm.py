"""The ALGORITHM of the time-parallel small-batch kernel (tp_eval_kernel), restated on the
numpy oracle's step function (oracle/tp_restatement.py), against the reference-style gradient (forward-mode
duals through the whole integration, problem.jl:154 as restated by oracle/oed_oracle.py).  No GPU: this pins
the three claims the kernel rests on — sensitivities affine in G given x, chunked evaluation of affine chains,
discrete adjoint over per-step Jacobians — independently of the CUDA code; the kernel itself is compared with
the oracle on the GPU (tests/test_gpu_parity.py::test_time_parallel_kernel_matches_the_other_gradient_kernels_and_the_oracle)."""
import numpy as np
import pytest

from dynamicoed.jl_b200 import models
from oracle import tp_restatement
from oracle.oed_oracle import Oracle

CASES = [("lv_fishing", models.lotka_volterra_fishing, 1, "rk4", 2),      # controls, FisherD
         ("lv_fishing", models.lotka_volterra_fishing, 3, "tsit5", 1),    # A criterion, six-stage tableau
         ("lv_ic", models.lotka_volterra_ic, 2, "rk4", 3)]                # unknown initial condition, FisherE


@pytest.mark.parametrize("name,ctor,crit,method,m", CASES)
@pytest.mark.parametrize("n_chunks", [1, 5, 16])
def test_time_parallel_algorithm_equals_the_forward_mode_gradient(name, ctor, crit, method, m, n_chunks):
    orc = Oracle(ctor(), crit, method, m)
    rng = np.random.Generator(np.random.Philox(7))
    des = rng.random((2, orc.n_var))
    des[:, -1] = 1e-3 + 0.9 * des[:, -1]
    oc, og, oF = orc.evaluate(des, grad=True)
    for b in range(2):
        obj, g, F, diag = tp_restatement.evaluate(orc, des[b], n_chunks=n_chunks)
        assert diag["G_chain_vs_sweep"] < 1e-12                              # claim 1 + 2: G from the affine chain
        assert abs(obj - oc[b]) <= 1e-11 * max(1.0, abs(oc[b]))
        assert np.max(np.abs(F - oF[b])) <= 1e-12 * np.max(np.abs(oF[b]))
        assert np.max(np.abs(g - og[b])) <= 1e-10 * np.max(np.abs(og[b])), (name, n_chunks)   # claim 3


def test_chunking_does_not_change_the_result_beyond_rounding():
    orc = Oracle(models.lotka_volterra_fishing(), 1, "rk4", 2)
    des = np.random.Generator(np.random.Philox(9)).random(orc.n_var)
    ref = tp_restatement.evaluate(orc, des, n_chunks=1)
    for nch in (2, 7, 16, 60):
        got = tp_restatement.evaluate(orc, des, n_chunks=nch)
        assert abs(got[0] - ref[0]) <= 1e-13 * abs(ref[0]) and np.max(np.abs(got[2] - ref[2])) <= 1e-13 * np.max(np.abs(ref[2]))
        assert np.max(np.abs(got[1] - ref[1])) <= 1e-12 * np.max(np.abs(ref[1]))


def test_time_parallel_algorithm_on_a_graded_grid():
    """Per-interval substep edges (step sizes differ from step to step): the per-step (t, h) travel along the
    batch axis of the oracle's step function."""
    N = 10
    edges = [np.array([0.0, 0.1, 0.35, 1.0])] + [np.array([0.0, 0.5, 1.0])] * (N - 1)
    orc = Oracle(models.lotka_volterra_ic(), 0, "rk4", edges)
    assert orc.N == N and orc.uniform is None
    des = np.random.Generator(np.random.Philox(11)).random((1, orc.n_var))
    oc, og, oF = orc.evaluate(des, grad=True)
    obj, g, F, diag = tp_restatement.evaluate(orc, des[0], n_chunks=4)
    assert diag["G_chain_vs_sweep"] < 1e-12 and abs(obj - oc[0]) <= 1e-11 * max(1.0, abs(oc[0]))
    assert np.max(np.abs(g - og[0])) <= 1e-10 * np.max(np.abs(og[0]))

