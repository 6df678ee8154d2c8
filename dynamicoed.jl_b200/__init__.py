"""dynamicoed.jl_b200 — B200-native batched evaluator for DynamicOED.jl's hot path.

Host-side mirror of the reference's public surface (/root/reference/src/DynamicOED.jl:31-42):
``OEDSystem``, the six criteria, ``OEDProblem``, ``get_timegrids``, ``get_initial_variables``,
plus ``OptimizationProblem`` — with objective and gradient evaluated by ``libdynoed_cuda.so``.
"""
from .symbolic import (Differential, Eq, Equation, Num, ODESystem, independent_variable,
                       observed_variables, parameters, variables)
from .augment import OEDSystem, dae_index_lowering, structural_simplify
from .criteria import (ACriterion, DCriterion, ECriterion, FisherACriterion, FisherDCriterion,
                       FisherECriterion, LogDCriterion)
from .discretize import ComponentVector, Timegrid
from .problem import (RK4, RODAS3, ROS2, ROS3P, ConstraintsSystem, OEDProblem, OptimizationProblem, Tsit5, graded_edges,
                      build_predictor, compute_information_gain, compute_local_information_gain,
                      continuous_predict, extract_fisher, geq, get_initial_variables, get_timegrids, leq, solve, states)

__all__ = ["OEDSystem", "FisherACriterion", "FisherDCriterion", "FisherECriterion", "ACriterion",
           "DCriterion", "ECriterion", "LogDCriterion", "OEDProblem", "get_timegrids", "get_initial_variables",
           "OptimizationProblem", "structural_simplify", "dae_index_lowering", "ODESystem",
           "variables", "parameters", "observed_variables", "Differential", "Eq", "states",
           "build_predictor", "continuous_predict", "compute_information_gain", "compute_local_information_gain",
           "extract_fisher", "ConstraintsSystem", "leq", "geq", "RK4", "Tsit5", "ROS2", "ROS3P", "RODAS3",
           "graded_edges", "Timegrid", "solve"]
