"""ctypes wrapper of oracle/_build/liboed_oracle.so (CPU ORACLE — test infrastructure only).

Takes its grid/layout tables from an `oed_oracle.Oracle` instance so that both oracles and the
GPU library see exactly the same step table (t_s, h_s) and indicator matrix.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboed_oracle.so")
MODEL_IDS = {"one_d": 0, "readme": 1, "lv_fishing": 2, "lv_ic": 3, "robertson": 4, "chem6": 5}


class _Desc(C.Structure):
    _fields_ = [("model", C.c_int32), ("crit", C.c_int32), ("method", C.c_int32), ("N", C.c_int32),
                ("n_var", C.c_int32), ("n_steps", C.c_int32),
                ("tgrid", C.c_void_p), ("first_step", C.c_void_p), ("step_t", C.c_void_p),
                ("step_h", C.c_void_p), ("ind", C.c_void_p), ("ctrl_off", C.c_void_p),
                ("w_off", C.c_void_p), ("ic_off", C.c_int32), ("tau_off", C.c_int32),
                ("theta", C.c_void_p), ("x0", C.c_void_p), ("G0", C.c_void_p)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return LIB


NATIVE_FLAGS = "-O3 -march=native -ffp-contract=fast"   # the timing arm's build (oracle/Makefile)


def _native_path():
    """The timing-arm build (-O3 -march=native, FMA contraction on) for THIS host: built where it runs,
    named after the host's ISA flags so that a copy built on another machine is never loaded."""
    import hashlib
    flags = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = line
                break
    except OSError:
        pass
    tag = hashlib.sha256(flags.encode()).hexdigest()[:10]
    path = os.path.join(_HERE, "_build", f"liboed_oracle_native.{tag}.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "oed_oracle.cpp")):
        subprocess.run(["make", "-s", "-C", _HERE, "native", f"NATIVE_TAG={tag}"], check=True)
    return path


_lib = None
_libs = {}


def lib(native: bool = False):
    """The strict-FP parity oracle (default) or, for the timed CPU arm only, the native-ISA build."""
    global _lib
    if native:
        if "native" not in _libs:
            _libs["native"] = _configure(C.CDLL(_native_path()))
        return _libs["native"]
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = _configure(C.CDLL(LIB))
    return _lib


def _configure(L):
    L.oracle_eval_batch.argtypes = [C.POINTER(_Desc), C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int]
    L.oracle_trajectory.argtypes = [C.POINTER(_Desc), C.c_void_p, C.c_void_p]
    L.oracle_criterion.restype = C.c_double
    L.oracle_criterion.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p]
    L.oracle_max_threads.restype = C.c_int
    return L


def uniform_step(tgrid, substeps):
    """The single step size of a grid that is uniform up to rounding, else None.

    Grid points such as k/10 are rounded decimals, so L_j/m scatters over a few ulp; a fixed-dt
    integrator run (`solve(...; adaptive=false, dt=dt)`) steps with ONE dt.  Rule (shared with
    dynoed_create): hbar = (t_N - t_0)/(N m); if every L_j/m is within 64 ulp of hbar, all steps
    use hbar."""
    N = len(tgrid) - 1
    hbar = (float(tgrid[N]) - float(tgrid[0])) / (float(N) * substeps)
    for j in range(N):
        h = (float(tgrid[j + 1]) - float(tgrid[j])) / substeps
        if not abs(h - hbar) <= 64.0 * 2.220446049250313e-16 * hbar:
            return None
    return hbar


def step_table(tgrid, substeps=None, edges=None):
    """(first_step[N+1], step_t, step_h) with the arithmetic the GPU library uses
    (dynamicoed.jl_b200/csrc/dynoed_api.cu, dynoed_create)."""
    first, st, sh = [0], [], []
    hbar = uniform_step(tgrid, substeps) if edges is None else None
    for j in range(len(tgrid) - 1):
        t0, L = float(tgrid[j]), float(tgrid[j + 1]) - float(tgrid[j])
        if edges is None:
            h = hbar if hbar is not None else L / substeps
            for s in range(substeps):
                st.append(t0 + s * h)
                sh.append(h)
        else:
            e = edges[j]
            for s in range(len(e) - 1):
                ts = t0 + float(e[s]) * L
                st.append(ts)
                sh.append((t0 + float(e[s + 1]) * L) - ts)
        first.append(len(st))
    return (np.array(first, dtype=np.int32), np.array(st, dtype=np.float64),
            np.array(sh, dtype=np.float64))


class CppOracle:
    def __init__(self, model: str, py_oracle, native: bool = False, own_constants: bool = False):
        """own_constants: use the model constants hard-coded in oed_oracle.cpp from the reference's test files
        (lv_fishing, one_d) instead of the ones `py_oracle` took from the product's models.py."""
        o = py_oracle
        self.o = o
        self._native = native
        self.model = MODEL_IDS[model]
        self.n_var, self.N = o.n_var, o.N
        self.nF = o.nF
        self.n_aug = o.n_aug
        tg = o.grid.tpoints
        first, st, sh = step_table(tg, o.uniform, None if o.uniform is not None else o.edges)
        lay = o.layout
        self._a = dict(
            tgrid=np.ascontiguousarray(tg), first=first, st=st, sh=sh,
            ind=np.ascontiguousarray((o.grid.indicators - 1).astype(np.int32)),
            ctrl=np.array(lay.control_offsets or [0], dtype=np.int32),
            w=np.array(lay.w_offsets, dtype=np.int32),
            theta=np.array([o.theta[p.name] for p in o.fixed_and_tunable], dtype=np.float64),
            x0=np.ascontiguousarray(o.x0, dtype=np.float64),
            G0=np.ascontiguousarray(o.G0.T.reshape(-1), dtype=np.float64))   # column-major
        d = _Desc()
        d.model, d.crit = self.model, o.crit
        d.method = {"rk4": 0, "tsit5": 1, "ros2": 2, "ros3p": 3, "rodas3": 4}[o.method]
        d.N, d.n_var, d.n_steps = o.N, o.n_var, len(st)
        a = self._a
        d.tgrid, d.first_step, d.step_t, d.step_h = (a["tgrid"].ctypes.data, a["first"].ctypes.data,
                                                     a["st"].ctypes.data, a["sh"].ctypes.data)
        d.ind, d.ctrl_off, d.w_off = a["ind"].ctypes.data, a["ctrl"].ctypes.data, a["w"].ctypes.data
        d.ic_off, d.tau_off = lay.ic_offset, lay.tau_offset
        d.theta, d.x0, d.G0 = a["theta"].ctypes.data, a["x0"].ctypes.data, a["G0"].ctypes.data
        if own_constants:
            assert model in ("lv_fishing", "one_d")
            d.theta = d.x0 = d.G0 = None
        self.d = d

    def set_criterion(self, crit: int):
        self.d.crit = int(crit)

    def evaluate(self, designs, grad=True, threads=0):
        designs = np.ascontiguousarray(np.atleast_2d(designs), dtype=np.float64)
        n = designs.shape[0]
        crit = np.empty(n)
        g = np.empty((n, self.n_var)) if grad else None
        f = np.empty((n, self.nF))
        rc = lib(self._native).oracle_eval_batch(C.byref(self.d), designs.ctypes.data, n, crit.ctypes.data,
                                     g.ctypes.data if grad else None, f.ctypes.data, threads)
        assert rc == 0
        return crit, g, f

    def trajectory(self, design):
        design = np.ascontiguousarray(design, dtype=np.float64)
        X = np.empty((self.N + 1, self.n_aug))
        rc = lib().oracle_trajectory(C.byref(self.d), design.ctypes.data, X.ctypes.data)
        assert rc == 0
        return X.T.copy()


def max_threads():
    return lib().oracle_max_threads()
