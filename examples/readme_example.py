#!/usr/bin/env python
"""The reference's README example (/root/reference/README.md:25-63), line for line, on the GPU
evaluator: x' = p x, observed x^2 ten times on (0, 1), FisherA criterion, at most 3 measurements."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynamicoed.jl_b200 as D                                             # noqa: E402
from dynamicoed.jl_b200 import (Differential, Eq, ODESystem, independent_variable,   # noqa: E402
                                observed_variables, parameters, variables)

# Define the differential equations
t = independent_variable("t")
x = variables("x", 1.0, description="State")
p = parameters("p[1]", -2.0, description="Fixed parameter", tunable=True)
obs = observed_variables("obs", description="Observed", measurement_rate=10)
Dt = Differential(t)

simple_system = ODESystem([Eq(Dt(x), p * x)], t, tspan=(0.0, 1.0), observed=[Eq(obs, x ** 2)],
                          name="simple_system")

oed = D.structural_simplify(D.OEDSystem(simple_system))

# Augment the original problem to an OED problem
oed_problem = D.OEDProblem(oed, D.FisherACriterion())

# Define a constraint system over the grid variables
optimization_variables = D.states(oed_problem)
constraint_equations = [D.leq(sum(optimization_variables.measurements["w₁"]), 3)]
constraint_set = D.ConstraintsSystem(constraint_equations, optimization_variables)

# Initialize the optimization problem (objective and gradient come from libdynoed_cuda.so)
optimization_problem = D.OptimizationProblem(oed_problem, None, constraints=constraint_set,
                                             integer_constraints=False)

# Solve for the optimal values of the observed variables
res = D.solve(optimization_problem)
w = res.u[:-1]
print("success", res.success, "objective", res.objective, "evaluations", res.evaluations)
print("w*  =", np.round(w, 4))
print("tau =", res.u[-1])
