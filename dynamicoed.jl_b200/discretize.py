"""Time-grid compiler and design-vector layout (host side, cold path).

Restates the behaviour of /root/reference/src/discretize.jl:
  * ``_generate_timegrid``          (:7-21)   per-variable grids from a rate (Real => step,
                                               Int => number of equal parts; tail interval kept)
  * ``Timegrid``                    (:43-76)  merged sorted-unique grid, intervals shorter than
                                               ``time_tolerance`` dropped, indicator matrix
  * ``get_variable_idx``            (:88-92), ``get_vars_from_grid`` (:94-99)
  * ``generate_initial_variables``  (:110-137), ``generate_variable_bounds`` (:139-173)
    -> the wire layout of one design: keys sorted alphabetically
       (/root/reference/src/DynamicOED.jl:19-21): controls | initial_conditions | measurements |
       regularization
Julia float ranges (``t0:dt:t1``) produce the double nearest to (a + k b)/den after rationalising
start/step/stop (Base twiceprecision.jl); `julia_float_range` restates that rule so that grid
points agree bit-for-bit with the reference's (e.g. 3*0.1 -> 0.3, not 0.30000000000000004).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

from .symbolic import (get_measurement_rate, getbounds, getdefault, is_measured)

_MAXINT_F32 = 16777216  # maxintfloat(Float32): bound used by Base.rat for Float64 inputs


def _rat(x: float):
    """Continued-fraction rationalisation used by Julia's float ranges (Base.rat)."""
    y = x
    a = d = 1
    b = c = 0
    m = float(_MAXINT_F32)
    while abs(y) <= m:
        f = int(math.trunc(y))
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > _MAXINT_F32:
            return c, d
        if b != 0 and a / b == x:
            break
        if y == 0:
            break
        y = 1.0 / y
    return a, b


def julia_float_range(start: float, step: float, stop: float) -> list[float]:
    """collect(start:step:stop) for Float64 arguments."""
    assert step != 0
    start, step, stop = float(start), float(step), float(stop)
    step_n, step_d = _rat(step)
    if step_d != 0 and step_n / step_d == step:
        start_n, start_d = _rat(start)
        stop_n, stop_d = _rat(stop)
        if start_d != 0 and stop_d != 0 and start_n / start_d == start and stop_n / stop_d == stop:
            den = math.lcm(start_d, step_d)
            m = 2.0 ** 53
            if den != 0 and abs(start * den) <= m and abs(step * den) <= m:
                sn = round(start * den)
                tn = round(step * den)
                num = den * stop_n - stop_d * sn
                dd = tn * stop_d
                q = abs(num) // abs(dd)
                q = q if (num >= 0) == (dd >= 0) else -q  # div truncates toward zero
                ln = max(0, int(q) + 1)
                lo, hi = start, stop + step / 2

                def between(a, x, b):
                    return a <= x <= b or b <= x <= a
                if between(start, start + (ln - 1) * step, hi) and \
                        not between(start, start + ln * step, stop):
                    return [float(Fraction(sn + k * tn, den)) for k in range(ln)]
    lf = (stop - start) / step
    if lf < 0:
        ln = 0
    elif lf == 0:
        ln = 1
    else:
        ln = int(round(lf)) + 1
        stop2 = start + (ln - 1) * step
        ln -= int(start < stop < stop2) + int(start > stop > stop2)
    fs, ft = Fraction(start), Fraction(step)
    return [float(fs + k * ft) for k in range(ln)]


def _generate_timegrid(rate, tspan) -> list[float]:
    """discretize.jl:7-21."""
    t0, tinf = float(tspan[0]), float(tspan[1])
    if isinstance(rate, (int, np.integer)) and not isinstance(rate, bool):
        dt = (tinf - t0) / int(rate)
    else:
        dt = float(rate)
    assert dt > 0, "Stepsize must be greater than 0."
    assert dt < tinf - t0, "Stepsize must be smaller than total time interval."
    pts = julia_float_range(t0, dt, tinf)
    if pts[-1] != tinf:
        pts.append(tinf)
    return pts


def generate_timegrid(x, tspan):
    assert is_measured(x), f"Provided variable {x} has no rate. No time grid can be created!"
    return _generate_timegrid(get_measurement_rate(x), tspan)


@dataclass
class Timegrid:
    """discretize.jl:32-76. ``indicators`` is 1-based like the reference (0 = no interval)."""
    variables: tuple          # variable names
    indicators: np.ndarray    # int [n_var, N]
    timegrids: list           # per variable: list of (t0, t1)
    timespans: list           # merged: list of (t0, t1)

    @classmethod
    def build(cls, tspan, *xs, time_tolerance: float = 1e-6):
        assert all(is_measured(x) for x in xs), \
            "Not all variables have rates associated with them. Can not generate time grid."
        grids = [generate_timegrid(x, tspan) for x in xs]
        complete = sorted(set(p for g in grids for p in g))
        completegrid = [(a, b) for a, b in zip(complete[:-1], complete[1:])
                        if abs(a - b) >= time_tolerance]
        timegrids = [[(a, b) for a, b in zip(g[:-1], g[1:]) if abs(a - b) >= time_tolerance]
                     for g in grids]
        ind = np.zeros((len(xs), len(completegrid)), dtype=np.int64)
        for i in range(len(xs)):
            for j, (t_start, t_stop) in enumerate(completegrid):
                for k, (t_min, t_max) in enumerate(timegrids[i], start=1):
                    if t_start >= t_min and t_stop <= t_max:
                        ind[i, j] = k
        return cls(tuple(x.name for x in xs), ind, timegrids, completegrid)

    @classmethod
    def from_system(cls, sys, tspan=None):
        tspan = tspan if tspan is not None else sys.tspan
        return cls.build(tspan, *sys.controls, *sys.w)

    def get_tspan(self, i: int):
        return self.timespans[i - 1]

    def _get_variable_idx(self, var: str):
        return self.variables.index(var) if var in self.variables else None

    def get_variable_idx(self, var: str, i: int) -> int:
        idx = self._get_variable_idx(var)
        if idx is None:
            return 1
        return int(self.indicators[idx, i - 1])

    @property
    def tpoints(self) -> np.ndarray:
        return np.array([self.timespans[0][0]] + [b for _, b in self.timespans], dtype=np.float64)


class ComponentVector:
    """A flat float64 vector with named, nested views (the subset of ComponentArrays.jl the
    reference uses: ``u.controls.u``, ``u.measurements.w₁``, ``u.regularization``)."""

    def __init__(self, data, axes):
        data = np.asarray(data)
        if data.dtype != object:
            data = np.asarray(data, dtype=np.float64)
        object.__setattr__(self, "data", data)
        object.__setattr__(self, "axes", axes)

    def __getattr__(self, key):
        axes = object.__getattribute__(self, "axes")
        if key in axes:
            sl = axes[key]
            if isinstance(sl, dict):
                return ComponentVector(self.data, sl)
            if isinstance(sl, int):
                return self.data[sl] if self.data.dtype == object else float(self.data[sl])
            return self.data[sl]
        raise AttributeError(key)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.__getattr__(key)
        return np.asarray(self)[key]

    def __setattr__(self, key, value):
        sl = self.axes[key]
        if isinstance(sl, dict):
            lo = min(_lo(s) for s in _leaves(sl)) if _leaves(sl) else 0
            hi = max(_hi(s) for s in _leaves(sl)) if _leaves(sl) else 0
            self.data[lo:hi] = value
        else:
            self.data[sl] = value

    def __array__(self, dtype=None, copy=None):
        lv = _leaves(self.axes)
        if not lv:
            return np.zeros(0)
        lo, hi = min(_lo(s) for s in lv), max(_hi(s) for s in lv)
        return np.asarray(self.data[lo:hi], dtype=dtype)

    def __len__(self):
        return len(np.asarray(self))

    def keys(self):
        return list(self.axes.keys())

    def copy(self):
        return ComponentVector(self.data.copy(), self.axes)

    def __repr__(self):
        return f"ComponentVector({ {k: np.asarray(getattr(self, k)) for k in self.axes} })"


def _leaves(ax):
    out = []
    for v in ax.values():
        out.extend(_leaves(v) if isinstance(v, dict) else [v])
    return out


def _lo(s):
    return s if isinstance(s, int) else s.start


def _hi(s):
    return s + 1 if isinstance(s, int) else s.stop


@dataclass
class DesignLayout:
    """Offsets of one design in the flat vector (discretize.jl:110-137)."""
    control_names: list
    control_offsets: list
    control_lengths: list
    ic_names: list
    ic_offset: int
    w_names: list
    w_offsets: list
    w_lengths: list
    tau_offset: int
    n_var: int

    def axes(self):
        ax = {
            "controls": {n: slice(o, o + l) for n, o, l in
                         zip(self.control_names, self.control_offsets, self.control_lengths)},
            "initial_conditions": {n: self.ic_offset + i for i, n in enumerate(self.ic_names)},
            "measurements": {n: slice(o, o + l) for n, o, l in
                             zip(self.w_names, self.w_offsets, self.w_lengths)},
            "regularization": self.tau_offset,
        }
        return ax

    def wrap(self, vec) -> ComponentVector:
        vec = np.asarray(vec, dtype=np.float64)
        assert vec.shape == (self.n_var,)
        return ComponentVector(vec, self.axes())


def design_layout(sys, tgrid: Timegrid) -> DesignLayout:
    off = 0
    c_names, c_off, c_len = [], [], []
    for c in sys.controls:
        n = len(tgrid.timegrids[tgrid._get_variable_idx(c.name)])
        c_names.append(c.name), c_off.append(off), c_len.append(n)
        off += n
    ic_names = [ic.name for ic in sys.x_ic]
    ic_offset = off
    off += len(ic_names)
    w_names, w_off, w_len = [], [], []
    for w in sys.w:
        n = len(tgrid.timegrids[tgrid._get_variable_idx(w.name)])
        w_names.append(w.name), w_off.append(off), w_len.append(n)
        off += n
    tau = off
    off += 1
    return DesignLayout(c_names, c_off, c_len, ic_names, ic_offset, w_names, w_off, w_len, tau, off)


def generate_initial_variables(sys, tgrid: Timegrid) -> ComponentVector:
    lay = design_layout(sys, tgrid)
    v = np.zeros(lay.n_var)
    for c, o, l in zip(sys.controls, lay.control_offsets, lay.control_lengths):
        v[o:o + l] = getdefault(c, 0.0)
    for i, ic in enumerate(sys.x_ic):
        v[lay.ic_offset + i] = getdefault(ic)
    for w, o, l in zip(sys.w, lay.w_offsets, lay.w_lengths):
        v[o:o + l] = getdefault(w, 0.0)
    v[lay.tau_offset] = 1.0   # discretize.jl:131
    return lay.wrap(v)


def generate_variable_bounds(sys, tgrid: Timegrid, lower: bool = False) -> ComponentVector:
    lay = design_layout(sys, tgrid)
    v = np.zeros(lay.n_var)
    pick = (lambda b: b[0]) if lower else (lambda b: b[1])
    for c, o, l in zip(sys.controls, lay.control_offsets, lay.control_lengths):
        v[o:o + l] = pick(getbounds(c))
    for i, ic in enumerate(sys.x_ic):
        v[lay.ic_offset + i] = pick(getbounds(ic))
    for w, o, l in zip(sys.w, lay.w_offsets, lay.w_lengths):
        v[o:o + l] = pick(getbounds(w))
    v[lay.tau_offset] = np.finfo(np.float64).eps if lower else 1.0   # discretize.jl:168
    return lay.wrap(v)
