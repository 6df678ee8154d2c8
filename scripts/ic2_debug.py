"""Debug: per-entry gradient error of the two initial-condition gradient kernels against the C++ oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
from oracle.cpp_oracle import CppOracle
from oracle.oed_oracle import Oracle
ROS = {"ros2": D.ROS2, "ros3p": D.ROS3P, "rodas3": D.RODAS3}
rng = np.random.Generator(np.random.Philox(5))
for method, msub in (("ros2", 3), ("ros2", 8), ("ros2", 24), ("ros3p", 3)):
    for crit_cls, crit_id in ((D.FisherACriterion, 0), (D.FisherDCriterion, 1), (D.ACriterion, 3)):
        for flag in ("", "1"):
            if flag: os.environ["DYNOED_NO_IC2"] = "1"
            else: os.environ.pop("DYNOED_NO_IC2", None)
            ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_ic()), crit_cls(), alg=ROS[method](msub)).context()
            rng = np.random.Generator(np.random.Philox(5)); des = rng.random((64, ctx.n_var)); des[:, -1] = 1e-3 + 0.9 * des[:, -1]; des[:, 0] = 0.1 + 0.9 * des[:, 0]
            c, g, F = ctx.evaluate(des)
            orc = CppOracle("lv_ic", Oracle(models.lotka_volterra_ic(), crit_id, method, msub))
            oc, og, oF = orc.evaluate(des, grad=True)
            err = np.abs(g - og) / np.max(np.abs(og), axis=1, keepdims=True)
            print(method, msub, crit_cls.__name__, "ic2" if not flag else "dual", "F", np.max(np.abs(F - oF)) / np.max(np.abs(oF)),
                  "grad err by column", np.array2string(err.max(axis=0), precision=1), "ic rel", np.max(np.abs(g[:, 0] - og[:, 0]) / np.abs(og[:, 0])))
            ctx.close()
