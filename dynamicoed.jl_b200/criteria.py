"""The six information criteria (/root/reference/src/fisher.jl:26-107) as callable objects.

``c(F)`` and ``c(F, tau)`` evaluate ``c(Symmetric(F + tau I))`` (fisher.jl:12-15) ON THE GPU via
``dynoed_criterion_batch``; inside a batched design evaluation the criterion is fused into the
integration kernel's epilogue and selected by ``c.id``.
"""
from __future__ import annotations

import numpy as np


def _symmetric_from_vector(x):
    """fisher.jl:3-10: packed upper triangle (x[j(j-1)/2+i], i<=j, 1-based) -> symmetric matrix."""
    x = np.asarray(x)
    n = int(np.sqrt(2 * x.shape[0] + 0.25) - 0.5)
    M = np.zeros((n, n), dtype=x.dtype)
    k = 0
    for j in range(n):
        for i in range(j + 1):
            M[i, j] = M[j, i] = x[k]
            k += 1
    return M


def _vector_from_symmetric(F):
    F = np.asarray(F, dtype=np.float64)
    n = F.shape[-1]
    # Symmetric(): the upper triangle is authoritative
    return np.stack([F[..., i, j] for j in range(n) for i in range(j + 1)], axis=-1)


class AbstractInformationCriterion:
    id = -1

    def __call__(self, F, tau=0.0, device: int = 0):
        from . import runtime
        F = np.asarray(F, dtype=np.float64)
        single = F.ndim == 2
        packed = _vector_from_symmetric(F if not single else F[None])
        val = runtime.criterion_batch(self.id, packed, tau, device=device)
        return float(val[0]) if single else val

    def gradient(self, F, tau=0.0, device: int = 0):
        """d c / d F as a symmetric matrix (GPU)."""
        from . import runtime
        F = np.asarray(F, dtype=np.float64)
        packed = _vector_from_symmetric(F[None])
        _, dL = runtime.criterion_batch(self.id, packed, tau, with_gradient=True, device=device)
        return _symmetric_from_vector(dL[0])

    def __repr__(self):
        return f"{type(self).__name__}()"


class FisherACriterion(AbstractInformationCriterion):
    """-tr(F)  (fisher.jl:26-30)"""
    id = 0


class FisherDCriterion(AbstractInformationCriterion):
    """-det(F)  (fisher.jl:41-45)"""
    id = 1


class FisherECriterion(AbstractInformationCriterion):
    """-min(eigvals(F))  (fisher.jl:56-60)"""
    id = 2


class ACriterion(AbstractInformationCriterion):
    """tr(inv(F))  (fisher.jl:71-76)"""
    id = 3


class DCriterion(AbstractInformationCriterion):
    """det(inv(F))  (fisher.jl:87-91)"""
    id = 4


class ECriterion(AbstractInformationCriterion):
    """max(eigvals(inv(F)))  (fisher.jl:102-107)"""
    id = 5


class LogDCriterion(AbstractInformationCriterion):
    """-log(det(F)): the log form of the D criterion.  NOT in the reference (fisher.jl has -det F and
    1/det F); BASELINE.json's north_star names "FisherD logdet F", so it is offered as an extension."""
    id = 6


ALL_CRITERIA = [FisherACriterion, FisherDCriterion, FisherECriterion, ACriterion, DCriterion,
                ECriterion]          # the reference's six; LogDCriterion is an extension
