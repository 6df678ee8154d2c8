"""CPU ORACLE (test infrastructure, not product code) — fixed-grid Rosenbrock step for the numpy
oracle (`oed_oracle.Oracle(method="ros2" | "ros3p")`), values only.

Restates how the reference treats the stiff configuration: a Rosenbrock-type method (Rodas5 there,
/root/reference/test/references/Rober.jl:29) applied to the FLAT augmented system
[x; vec(G); F_triu] of /root/reference/src/augment.jl:240 with its dense Jacobian.  The Jacobian
is obtained by pushing one-direction dual numbers through `Oracle.rhs` (one column per
direction), the linear systems are solved by LAPACK (`numpy.linalg.solve`, batched).  Independent
of the product's emitter (symbolic second derivatives, block-triangular solves) and of the C++
oracle's nested-dual implementation.

Implementation form (Hairer/Wanner IV.7):
    (1/(h gamma) I - J) U_i = g(y + sum_{j<i} a_ij U_j) + sum_{j<i} (c_ij/h) U_j ,  y+ = y + sum m_i U_i
ROS2: Verwer, Spee, Blom, Hundsdorfer (1999); ROS3P: Lang, Verwer (2001); RODAS3: Sandu et al. (1997).
"""
from __future__ import annotations

import numpy as np

_G2 = 1.7071067811865475244
TABLEAUS = {
    "ros2": (_G2, [[0, 0], [1.0 / _G2, 0]], [[0, 0], [-2.0 / _G2, 0]], [1.5 / _G2, 0.5 / _G2]),
    "ros3p": (7.886751345948129e-01,
              [[0, 0, 0], [1.267949192431123, 0, 0], [1.267949192431123, 0, 0]],
              [[0, 0, 0], [-1.607695154586736, 0, 0], [-3.464101615137755, -1.732050807568877, 0]],
              [2.0, 5.773502691896258e-01, 4.226497308103742e-01]),
    "rodas3": (0.5, [[0, 0, 0, 0], [0, 0, 0, 0], [2.0, 0, 0, 0], [2.0, 0, 1.0, 0]],
               [[0, 0, 0, 0], [4.0, 0, 0, 0], [1.0, -1.0, 0, 0], [1.0, -1.0, -8.0 / 3.0, 0]],
               [2.0, 0.0, 1.0, 1.0]),
}


class Du:
    """value + eps * derivative, components are numpy arrays [B] (or scalars)."""
    __slots__ = ("v", "d")
    __array_priority__ = 2000

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def _lift(o):
        return o if isinstance(o, Du) else Du(o, 0.0)

    def __add__(self, o):
        o = Du._lift(o)
        return Du(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __sub__(self, o):
        o = Du._lift(o)
        return Du(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        o = Du._lift(o)
        return Du(o.v - self.v, o.d - self.d)

    def __neg__(self):
        return Du(-self.v, -self.d)

    def __mul__(self, o):
        o = Du._lift(o)
        return Du(self.v * o.v, self.d * o.v + self.v * o.d)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Du._lift(o)
        q = self.v / o.v
        return Du(q, (self.d - q * o.d) / o.v)

    def __rtruediv__(self, o):
        return Du._lift(o) / self

    def __pow__(self, e):
        if isinstance(e, (int, np.integer)) and e >= 0:
            out = Du(1.0, 0.0)
            for _ in range(int(e)):
                out = out * self
            return out
        return Du(self.v ** e, e * self.v ** (e - 1) * self.d)


def step(orc, t, y, h, u, w):
    """One Rosenbrock step of the flat augmented state `y` (list of [B] arrays)."""
    gamma, A, C, m = TABLEAUS[orc.method]
    S, na = len(m), len(y)
    B = np.broadcast(*[np.asarray(v) for v in y]).shape[0] if np.ndim(y[0]) else 1
    yv = [np.broadcast_to(np.asarray(v, dtype=np.float64), (B,)) for v in y]
    J = np.zeros((B, na, na))
    for c in range(na):
        yd = [Du(yv[a], np.ones(B) if a == c else np.zeros(B)) for a in range(na)]
        gd, _ = orc.rhs(t, yd, list(u), list(w))
        for a in range(na):
            ga = gd[a]
            J[:, a, c] = ga.d if isinstance(ga, Du) else 0.0
    W = np.eye(na)[None] / (h * gamma) - J
    U = []
    for s in range(S):
        Y = list(yv)
        for j in range(s):
            if A[s][j] != 0.0:
                Y = [Ya + A[s][j] * U[j][:, a] for a, Ya in enumerate(Y)]
        g, _ = orc.rhs(t, Y, list(u), list(w))
        r = np.stack([np.broadcast_to(np.asarray(ga, dtype=np.float64), (B,)) for ga in g], axis=1)
        for j in range(s):
            if C[s][j] != 0.0:
                r = r + (C[s][j] / h) * U[j]
        U.append(np.linalg.solve(W, r[..., None])[..., 0])
    out = np.stack(yv, axis=1)
    for s in range(S):
        out = out + m[s] * U[s]
    return [out[:, a] for a in range(na)]
