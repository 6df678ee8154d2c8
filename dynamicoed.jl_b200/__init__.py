"""dynamicoed.jl_b200 — B200-native batched evaluator for DynamicOED.jl's hot path."""
from .symbolic import (Differential, Eq, Equation, ODESystem, Num, independent_variable,
                       observed_variables, parameters, variables)
from .augment import OEDSystem, structural_simplify, dae_index_lowering
from .discretize import Timegrid
