"""Model definitions used by the reference's own tests/docs, written with the symbolic front-end.

Each function returns the *un-augmented* ``ODESystem`` exactly as the reference declares it
(file:line cited per model); ``OEDSystem`` + ``structural_simplify`` then produce the augmented
system that the CUDA emitter prints.  ``BUILTIN`` lists the models compiled into
``libdynoed_cuda.so``.
"""
from __future__ import annotations

import sympy as sp

from .symbolic import (Differential, Eq, ODESystem, independent_variable, observed_variables,
                       parameters, variables)


def one_d():
    """/root/reference/test/references/1D.jl:14-24 — x' = p x, y = x, 10 measurements on (0,1)."""
    t = independent_variable("t")
    x = variables("x", 1.0, description="State")
    p = parameters("p[1]", -2.0, description="Fixed parameter", tunable=True)
    obs = observed_variables("obs", description="Observed", measurement_rate=10)
    D = Differential(t)
    return ODESystem([Eq(D(x), p * x)], t, tspan=(0.0, 1.0), observed=[Eq(obs, x)],
                     name="simple_system")


def readme():
    """/root/reference/README.md:31-40 — x' = p x, y = x^2, 10 measurements on (0,1)."""
    t = independent_variable("t")
    x = variables("x", 1.0, description="State")
    p = parameters("p[1]", -2.0, description="Fixed parameter", tunable=True)
    obs = observed_variables("obs", description="Observed", measurement_rate=10)
    D = Differential(t)
    return ODESystem([Eq(D(x), p * x)], t, tspan=(0.0, 1.0), observed=[Eq(obs, x ** 2)],
                     name="simple_system")


def lotka_volterra_fishing(integer_control: bool = False):
    """/root/reference/test/references/LotkaVolterra.jl:10-36 (relaxed) and :158-187 (u::Int)."""
    t = independent_variable("t")
    x, y = variables("x y", [0.5, 0.7])
    p = parameters("p[1] p[2] p[3] p[4]", [1.0, 1.0, 0.4, 0.6], tunable=False,
                   description="Fixed parameters")
    c = parameters("c[1] c[2]", [1.0, 1.0], tunable=True, description="Uncertain parameters")
    u = parameters("u", 0.0, measurement_rate=0.3, input=True, bounds=(0.0, 1.0),
                   integer=integer_control, description="Binary control variable")
    obs = observed_variables("obs[1] obs[2]", measurement_rate=0.1)
    D = Differential(t)
    return ODESystem([Eq(D(x), p[0] * x - c[0] * x * y - p[2] * x * u),
                      Eq(D(y), -p[1] * y + c[1] * x * y - p[3] * y * u)],
                     t, tspan=(0.0, 3.0), observed=[Eq(obs[0], x), Eq(obs[1], y)],
                     name="lotka_volterra")


def lotka_volterra_ic():
    """/root/reference/test/references/LotkaVolterra.jl:305-330 — unknown initial condition x0."""
    t = independent_variable("t")
    x = variables("x", 0.49, tunable=True, bounds=(0.1, 1.0), description="Biomass Prey")
    y = variables("y", 0.7, tunable=False, description="Biomass Predator")
    p = parameters("p[1] p[2] p[3]", [1.0, 1.0, 1.0], tunable=False)
    c = parameters("c[1]", 1.0, tunable=True)
    obs = observed_variables("obs[1]", measurement_rate=10)
    D = Differential(t)
    return ODESystem([Eq(D(x), p[0] * x - c * x * y),
                      Eq(D(y), -p[1] * y + p[2] * x * y)],
                     t, tspan=(0.0, 5.0), observed=[Eq(obs, y + x)], name="lotka_volterra")


def robertson():
    """/root/reference/test/references/Rober.jl:9-21 — Robertson DAE, k1 tunable, obs y1."""
    t = independent_variable("t")
    y1, y2, y3 = variables("y₁ y₂ y₃", [1.0, 0.0, 0.0])
    obs = observed_variables("obs", measurement_rate=10)
    k1 = parameters("k₁", 0.04, tunable=True)
    k2, k3 = parameters("k₂ k₃", [3e7, 1e4])
    D = Differential(t)
    eqs = [Eq(D(y1), -k1 * y1 + k3 * y2 * y3),
           Eq(D(y2), k1 * y1 - k3 * y2 * y3 - k2 * y2 ** 2),
           Eq(0, y1 + y2 + y3 - 1)]
    return ODESystem(eqs, t, tspan=(0.0, 100.0), observed=[Eq(obs, y1)], name="roberdae")


def chem6():
    """The builder-defined stiff instance of BASELINE.json config 4 ("stiff chemical kinetics with
    unknown initial conditions, ~6 states, 12 sensitivities, FisherE, Rosenbrock"): it is NOT in
    the reference (sizes table).  Robertson's autocatalytic core
    (/root/reference/test/references/Rober.jl:9-21: A -> B, B + C -> A + C, 2B -> B + C) feeding a
    slow cycle C + D -> E -> F -> D.  The two unknown initial conditions A(0), D(0) are the
    LEADING states (quirk Q3, /root/reference/src/augment.jl:186); observed: E and C."""
    t = independent_variable("t")
    A = variables("A", 1.0, tunable=True, bounds=(0.5, 1.5))
    Dd = variables("D", 0.5, tunable=True, bounds=(0.1, 1.0))
    B, Cc, E, Fs = variables("B C E Fs", [0.0, 0.0, 0.0, 0.0])
    k1, k2, k3, k4, k5, k6 = parameters("k₁ k₂ k₃ k₄ k₅ k₆", [0.04, 3e7, 1e4, 50.0, 2.0, 0.5],
                                        tunable=False)
    obs = observed_variables("obs[1] obs[2]", measurement_rate=10)
    D = Differential(t)
    eqs = [Eq(D(A), -k1 * A + k3 * B * Cc),
           Eq(D(Dd), -k4 * Cc * Dd + k6 * Fs),
           Eq(D(B), k1 * A - k3 * B * Cc - k2 * B ** 2),
           Eq(D(Cc), k2 * B ** 2 - k4 * Cc * Dd),
           Eq(D(E), k4 * Cc * Dd - k5 * E),
           Eq(D(Fs), k5 * E - k6 * Fs)]
    return ODESystem(eqs, t, tspan=(0.0, 50.0), observed=[Eq(obs[0], E), Eq(obs[1], Cc)],
                     name="chem6")


def mtk_simple(n_obs: int = 1, control: bool = False):
    """/root/reference/test/mtk_extensions.jl:6-21 and /root/reference/test/discretize.jl:6-46 —
    x' = p x (+ c), y1 = x (rate 1.0), y2 = sqrt(x) (rate 0.3), c rate 0.1, on (0,2)."""
    t = independent_variable("t")
    x = variables("x", 2.0, tunable=False, bounds=(0.1, 5.0))
    p = parameters("p[1]", -2.0, tunable=True)
    y1 = observed_variables("y₁", measurement_rate=1.0)
    y2 = observed_variables("y₂", measurement_rate=0.3)
    c = parameters("c", 0.0, input=True, measurement_rate=0.1)
    D = Differential(t)
    rhs = p * x + (c if control else 0)
    obs = [Eq(y1, x)] + ([Eq(y2, sp.sqrt(x))] if n_obs > 1 else [])
    return ODESystem([Eq(D(x), rhs)], t, tspan=(0.0, 2.0), observed=obs, name="simple_system")


BUILTIN = {
    "one_d": one_d,
    "readme": readme,
    "lv_fishing": lotka_volterra_fishing,
    "lv_ic": lotka_volterra_ic,
    "robertson": robertson,
    "chem6": chem6,
}
