"""ctypes binding of ``libdynoed_cuda.so`` (C ABI in ``include/dynoed.h``).

This is the only way the package computes anything: there is no CPU fallback.  If the shared
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``) importing
this module raises; if no CUDA device is present every compute call raises `DynoedError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DYNOED_LIB: another build of the same library (measurement variants, scripts/build_variants.py)
LIB_PATH = os.environ.get("DYNOED_LIB") or os.path.join(_HERE, "libdynoed_cuda.so")

RK4, TSIT5, ROS2, ROS3P, RODAS3 = 0, 1, 2, 3, 4
GRAD_NO_CONTROLS = 1    # eval flag: weights / tau / IC gradient only (forward sweep), control entries NaN
OVERLAP_OK = 2          # eval flag: inputs are not produced by work enqueued after the previous launch (see dynoed.h)
METHODS = {"rk4": RK4, "RK4": RK4, "tsit5": TSIT5, "Tsit5": TSIT5, "ros2": ROS2, "ROS2": ROS2,
           "ros3p": ROS3P, "ROS3P": ROS3P, "rodas3": RODAS3, "RODAS3": RODAS3}


class DynoedError(RuntimeError):
    pass


class ModelInfo(C.Structure):
    _fields_ = [("nx", C.c_int32), ("np", C.c_int32), ("n_obs", C.c_int32), ("n_ctrl", C.c_int32),
                ("n_ic", C.c_int32), ("n_theta", C.c_int32), ("nz", C.c_int32), ("nF", C.c_int32),
                ("autonomous", C.c_int32), ("ic_state", C.c_int32 * 8),
                ("flops_primal", C.c_int32), ("flops_primal_z", C.c_int32),
                ("flops_adjoint", C.c_int32), ("flops_consts", C.c_int32), ("flops_ros_f", C.c_int32),
                ("x0", C.c_double * 8), ("G0", C.c_double * 64), ("theta", C.c_double * 48)]


class ProblemDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("model", C.c_char_p), ("criterion", C.c_int32),
                ("method", C.c_int32), ("n_intervals", C.c_int32),
                ("tgrid", C.POINTER(C.c_double)), ("substeps", C.c_int32),
                ("edge_count", C.POINTER(C.c_int32)), ("step_edges", C.POINTER(C.c_double)),
                ("indicators", C.POINTER(C.c_int32)), ("ctrl_offsets", C.POINTER(C.c_int32)),
                ("w_offsets", C.POINTER(C.c_int32)), ("ic_offset", C.c_int32),
                ("tau_offset", C.c_int32), ("n_var", C.c_int32),
                ("theta", C.POINTER(C.c_double)), ("x0", C.POINTER(C.c_double)),
                ("G0", C.POINTER(C.c_double)), ("tile", C.c_int32), ("max_batch_hint", C.c_int32)]


class KernelStats(C.Structure):
    _fields_ = [("regs_per_thread", C.c_int32), ("static_smem", C.c_int32),
                ("dynamic_smem", C.c_int32), ("blocks_per_sm", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32), ("sm_count", C.c_int32), ("launches", C.c_int64),
                ("scratch_bytes", C.c_int64), ("n_steps", C.c_int32), ("stages", C.c_int32),
                ("flops_per_design_eval", C.c_double), ("flops_per_design_grad", C.c_double),
                ("bytes_per_design", C.c_double), ("overlapped_launches", C.c_int64),
                ("small_batch_launches", C.c_int64), ("small_batch_max", C.c_int32), ("time_parallel_max", C.c_int32)]


EXPORTS = ["dynoed_abi_version", "dynoed_model_count", "dynoed_model_name", "dynoed_model_query",
           "dynoed_create", "dynoed_destroy", "dynoed_set_criterion", "dynoed_set_stream",
           "dynoed_own_stream",
           "dynoed_eval_batch", "dynoed_eval_batch_async", "dynoed_eval_batch_host_async", "dynoed_wait",
           "dynoed_sync", "dynoed_argmin",
           "dynoed_argmin_async", "dynoed_argmin_record_async", "dynoed_argmin_history_async",
           "dynoed_argmin_gathered", "dynoed_set_aux_stream", "dynoed_pack_winner", "dynoed_select_winner",
           "dynoed_guard_stats",
           "dynoed_set_profiling", "dynoed_kernel_time",
           "dynoed_trajectory", "dynoed_information_gain", "dynoed_fim_basis", "dynoed_hessian_batch",
           "dynoed_eval_fixed_controls", "dynoed_criterion_batch", "dynoed_fp64_peak", "dynoed_fp64_microbench",
           "dynoed_kernel_stats_query", "dynoed_host_alloc", "dynoed_host_free",
           "dynoed_last_error"]

_libs = {}


def lib(path: str | None = None):
    """The configured CDLL for ``path`` (default: the in-tree ``libdynoed_cuda.so``)."""
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise DynoedError(f"{path} is missing: build it with __graft_entry__.build(); "
                              "there is no CPU fallback")
        L = C.CDLL(path)
        vp, dp, i64 = C.c_void_p, C.c_void_p, C.c_int64
        L.dynoed_abi_version.restype = C.c_int
        L.dynoed_model_count.restype = C.c_int
        L.dynoed_model_name.restype = C.c_char_p
        L.dynoed_model_name.argtypes = [C.c_int]
        L.dynoed_model_query.argtypes = [C.c_char_p, C.POINTER(ModelInfo)]
        L.dynoed_create.argtypes = [C.POINTER(ProblemDesc), C.c_int, C.POINTER(vp)]
        L.dynoed_destroy.argtypes = [vp]
        L.dynoed_destroy.restype = None
        L.dynoed_set_criterion.argtypes = [vp, C.c_int]
        L.dynoed_set_stream.argtypes = [vp, vp]
        L.dynoed_own_stream.argtypes = [vp]
        L.dynoed_eval_batch.argtypes = [vp, dp, i64, dp, dp, dp, C.c_uint32]
        L.dynoed_eval_batch_async.argtypes = [vp, dp, i64, dp, dp, dp, C.c_uint32]
        L.dynoed_eval_batch_host_async.argtypes = [vp, dp, i64, dp, dp, dp, C.c_uint32, C.POINTER(C.c_int64)]
        L.dynoed_wait.argtypes = [vp, i64]
        L.dynoed_sync.argtypes = [vp]
        L.dynoed_argmin.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.dynoed_argmin_async.argtypes = [vp, dp, dp]
        L.dynoed_argmin_record_async.argtypes = [vp, dp]
        L.dynoed_argmin_history_async.argtypes = [vp, C.c_int, dp]
        L.dynoed_argmin_gathered.argtypes = [vp, dp, C.c_int, C.c_int, i64, dp]
        L.dynoed_set_aux_stream.argtypes = [vp, vp]
        L.dynoed_pack_winner.argtypes = [vp, C.c_int, dp, dp, i64, i64, i64, dp]
        L.dynoed_select_winner.argtypes = [vp, dp, C.c_int, dp]
        L.dynoed_guard_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.dynoed_set_profiling.argtypes = [vp, C.c_int]
        L.dynoed_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.dynoed_trajectory.argtypes = [vp, dp, i64, dp]
        L.dynoed_information_gain.argtypes = [vp, dp, i64, dp, dp, dp]
        L.dynoed_fim_basis.argtypes = [vp, dp, dp]
        L.dynoed_hessian_batch.argtypes = [vp, dp, C.c_int64, C.c_double, dp, dp, dp]
        L.dynoed_eval_fixed_controls.argtypes = [vp, dp, dp, i64, dp, dp, dp]
        L.dynoed_criterion_batch.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, i64, dp, dp]
        L.dynoed_fp64_peak.argtypes = [C.c_int, vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.dynoed_fp64_microbench.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.dynoed_kernel_stats_query.argtypes = [vp, C.c_int, C.POINTER(KernelStats)]
        L.dynoed_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.dynoed_host_free.argtypes = [vp]
        L.dynoed_last_error.restype = C.c_char_p
        L.dynoed_last_error.argtypes = [vp]
        if L.dynoed_abi_version() != 1:
            raise DynoedError("libdynoed_cuda.so ABI version mismatch")
        _libs[path] = L
    return _libs[path]


def _check(rc, ctx=None, L=None):
    if rc != 0:
        msg = (L or lib()).dynoed_last_error(ctx)
        raise DynoedError(f"dynoed error {rc}: {msg.decode() if msg else '?'}")


def model_names(path=None):
    L = lib(path)
    return [L.dynoed_model_name(i).decode() for i in range(L.dynoed_model_count())]


def model_info(name: str, path=None) -> ModelInfo:
    mi = ModelInfo()
    _check(lib(path).dynoed_model_query(name.encode(), C.byref(mi)), None, lib(path))
    return mi


def _ptr(a):
    """Device pointer of a torch CUDA tensor, host pointer of a numpy array / torch CPU tensor."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous() and str(a.dtype) == "torch.float64"
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(type(a))


class Context:
    """One OED problem bound to one GPU (wraps ``dynoed_ctx``)."""

    def __init__(self, model: str, *, criterion: int, method: int, tgrid, indicators, ctrl_offsets,
                 w_offsets, ic_offset: int, tau_offset: int, n_var: int, substeps: int = 4,
                 step_edges=None, theta=None, x0=None, G0=None, device: int = 0, tile: int = 0,
                 library_path: str | None = None):
        L = self._L = lib(library_path)
        self._keep = []

        def arr(x, dt):
            if x is None:
                return None
            a = np.ascontiguousarray(np.asarray(x, dtype=dt))
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double if dt == np.float64 else C.c_int32))
        tgrid = np.asarray(tgrid, dtype=np.float64)
        d = ProblemDesc()
        d.struct_size = C.sizeof(ProblemDesc)
        d.model = model.encode()
        d.criterion = int(criterion)
        d.method = int(method)
        d.n_intervals = len(tgrid) - 1
        d.tgrid = arr(tgrid, np.float64)
        if step_edges is not None:
            d.substeps = 0
            d.edge_count = arr([len(e) for e in step_edges], np.int32)
            d.step_edges = arr(np.concatenate([np.asarray(e, dtype=np.float64) for e in step_edges]),
                               np.float64)
        else:
            d.substeps = int(substeps)
        d.indicators = arr(np.asarray(indicators, dtype=np.int32).reshape(-1), np.int32)
        d.ctrl_offsets = arr(list(ctrl_offsets) or [0], np.int32)
        d.w_offsets = arr(list(w_offsets), np.int32)
        d.ic_offset = int(ic_offset)
        d.tau_offset = int(tau_offset)
        d.n_var = int(n_var)
        d.theta = arr(theta, np.float64)
        d.x0 = arr(x0, np.float64)
        d.G0 = arr(G0, np.float64)
        d.tile = int(tile)
        self.n_var = int(n_var)
        self.n_intervals = d.n_intervals
        self.info = model_info(model, library_path)
        self.device = device
        h = C.c_void_p()
        _check(L.dynoed_create(C.byref(d), device, C.byref(h)), None, L)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.dynoed_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_criterion(self, criterion: int):
        _check(self._L.dynoed_set_criterion(self._h, int(criterion)), self._h, self._L)

    def set_stream(self, stream_ptr: int):
        """Run on the given cudaStream_t (0 = legacy default stream, e.g. torch's default)."""
        _check(self._L.dynoed_set_stream(self._h, C.c_void_p(int(stream_ptr))), self._h, self._L)
        self._bound_stream = int(stream_ptr)

    def own_stream(self):
        _check(self._L.dynoed_own_stream(self._h), self._h, self._L)
        self._bound_stream = None

    def eval_batch(self, designs, crit, grad=None, fim=None, n=None, async_=False, flags=0):
        """designs [n, n_var], crit [n], grad [n, n_var] | None, fim [n, nF] | None — numpy arrays
        (host path) or torch CUDA tensors (device path), float64, contiguous."""
        n = int(designs.shape[0]) if n is None else int(n)
        if getattr(designs, "is_cuda", False):
            self.bind_torch_stream(designs.device)
        fn = self._L.dynoed_eval_batch_async if async_ else self._L.dynoed_eval_batch
        _check(fn(self._h, _ptr(designs), n, _ptr(crit), _ptr(grad), _ptr(fim), int(flags)), self._h, self._L)

    def bind_torch_stream(self, device=None):
        """Order this context with torch: run on torch's CURRENT stream of `device`.  CUDA tensors handed to
        `eval_batch` were produced, and their results will be consumed, on that stream; a context left on its
        own non-blocking stream would race with both.  (No-op when the stream is unchanged.)"""
        import torch
        ptr = torch.cuda.current_stream(device).cuda_stream
        if ptr != getattr(self, "_bound_stream", None):
            self.set_stream(ptr)


    def eval_batch_host_async(self, designs, crit, grad=None, fim=None, flags=0) -> int:
        """Pipelined host-buffer evaluation (page-locked numpy arrays): returns a call id for `wait`."""
        cid = C.c_int64()
        _check(self._L.dynoed_eval_batch_host_async(self._h, _ptr(designs), int(designs.shape[0]), _ptr(crit),
                                                    _ptr(grad), _ptr(fim), int(flags), C.byref(cid)), self._h, self._L)
        return int(cid.value)

    def wait(self, call_id: int):
        _check(self._L.dynoed_wait(self._h, int(call_id)), self._h, self._L)

    def sync(self):
        _check(self._L.dynoed_sync(self._h), self._h, self._L)

    def argmin(self):
        i, v = C.c_int64(), C.c_double()
        _check(self._L.dynoed_argmin(self._h, C.byref(i), C.byref(v)), self._h, self._L)
        return int(i.value), float(v.value)

    def argmin_async(self, d_val, d_idx):
        """Copy the (value, index) record into 1-element CUDA tensors (float64 / int64)."""
        _check(self._L.dynoed_argmin_async(self._h, d_val.data_ptr(), d_idx.data_ptr()), self._h, self._L)

    def argmin_record_async(self, d_record):
        """Copy the 16-byte {value, local index} record into a 2-element int64 CUDA tensor."""
        _check(self._L.dynoed_argmin_record_async(self._h, d_record.data_ptr()), self._h, self._L)

    def argmin_history_async(self, count: int, d_records):
        """Records of the last `count` batches (oldest first) into a [count, 2] int64 CUDA tensor."""
        _check(self._L.dynoed_argmin_history_async(self._h, int(count), d_records.data_ptr()), self._h, self._L)

    def argmin_gathered(self, d_records, world: int, shard_size: int, d_out, groups: int = 1):
        """Arg-min over all-gathered records ([world, groups, 2] int64 CUDA tensor) into d_out
        ([groups, 3] int64: value bits, global index, owner rank), on the ctx stream."""
        _check(self._L.dynoed_argmin_gathered(self._h, d_records.data_ptr(), int(world), int(groups),
                                              int(shard_size), d_out.data_ptr()), self._h, self._L)

    def set_aux_stream(self, stream_ptr: int):
        """Selection helpers run on this cudaStream_t (0 = back to the ctx stream); see dynoed.h."""
        _check(self._L.dynoed_set_aux_stream(self._h, C.c_void_p(int(stream_ptr)) if stream_ptr else None),
               self._h, self._L)

    def pack_winner(self, nbatches: int, designs, grad, batch_stride: int, index_base: int, d_payload, n_local=None):
        """[value, global index, design row, gradient row] of the best candidate among the last
        `nbatches` batches, written to the CUDA tensor d_payload (2 + 2 n_var float64)."""
        n_local = int(designs.shape[0]) if n_local is None else int(n_local)
        _check(self._L.dynoed_pack_winner(self._h, int(nbatches), _ptr(designs) if n_local else None,
                                          _ptr(grad) if (grad is not None and n_local) else None, n_local,
                                          int(batch_stride), int(index_base), _ptr(d_payload)), self._h, self._L)

    def select_winner(self, d_gathered, world: int, d_out):
        """Global winner of all-gathered payloads [world, 2 + 2 n_var] -> d_out [2 + 2 n_var + 1] (owner last)."""
        _check(self._L.dynoed_select_winner(self._h, _ptr(d_gathered), int(world), _ptr(d_out)), self._h, self._L)

    def guard_stats(self):
        w, t = C.c_int64(), C.c_int64()
        _check(self._L.dynoed_guard_stats(self._h, C.byref(w), C.byref(t)), self._h, self._L)
        return int(w.value), int(t.value)

    def set_profiling(self, on: bool):
        _check(self._L.dynoed_set_profiling(self._h, int(bool(on))), self._h, self._L)

    def kernel_time(self):
        ms, cnt = C.c_double(), C.c_int64()
        _check(self._L.dynoed_kernel_time(self._h, C.byref(ms), C.byref(cnt)), self._h, self._L)
        return float(ms.value), int(cnt.value)

    def trajectory(self, designs):
        designs = np.ascontiguousarray(np.atleast_2d(designs), dtype=np.float64)
        n = designs.shape[0]
        X = np.empty((n, self.n_intervals + 1, self.info.nz + self.info.nF))
        _check(self._L.dynoed_trajectory(self._h, _ptr(designs), n, _ptr(X)), self._h, self._L)
        return X

    def information_gain(self, designs):
        """(F [n, nF], P [n, N+1, n_obs, np, np], Pi same shape) — utilities.jl:47-70 at the grid points."""
        designs = np.ascontiguousarray(np.atleast_2d(designs), dtype=np.float64)
        n, mi = designs.shape[0], self.info
        F = np.empty((n, mi.nF))
        P = np.empty((n, self.n_intervals + 1, mi.n_obs, mi.np, mi.np))
        Pi = np.empty_like(P)
        _check(self._L.dynoed_information_gain(self._h, _ptr(designs), n, _ptr(F), _ptr(P), _ptr(Pi)),
               self._h, self._L)
        return F, P, Pi

    def fim_basis(self, design):
        """basis [n_var, nF]: F(t_f) per unit measurement weight, controls / ICs taken from `design`."""
        design = np.ascontiguousarray(np.asarray(design, dtype=np.float64).reshape(self.n_var))
        basis = np.empty((self.n_var, self.info.nF))
        _check(self._L.dynoed_fim_basis(self._h, _ptr(design), _ptr(basis)), self._h, self._L)
        return basis

    def hessian_batch(self, designs, hess, crit=None, grad=None, rel_step=0.0, n=None):
        """Hessians of the objective (central differences of the exact gradient, evaluated as batches
        of the gradient kernel); numpy arrays or torch CUDA tensors, like `eval_batch`."""
        n = int(designs.shape[0]) if n is None else int(n)
        _check(self._L.dynoed_hessian_batch(self._h, _ptr(designs), n, float(rel_step), _ptr(crit),
                                            _ptr(grad), _ptr(hess)), self._h, self._L)

    def hessian(self, design, rel_step=0.0):
        """(objective, gradient, Hessian) at one or more host designs."""
        d = np.ascontiguousarray(np.atleast_2d(design), dtype=np.float64)
        n = d.shape[0]
        c, g, H = np.empty(n), np.empty((n, self.n_var)), np.empty((n, self.n_var, self.n_var))
        self.hessian_batch(d, H, c, g, rel_step)
        return (c[0], g[0], H[0]) if np.ndim(design) == 1 else (c, g, H)

    def eval_fixed_controls(self, basis, designs, crit, grad=None, fim=None, n=None):
        """Objective / gradient / F for measurement designs that share the controls of `basis`
        (numpy arrays or torch CUDA tensors, like `eval_batch`)."""
        n = int(designs.shape[0]) if n is None else int(n)
        _check(self._L.dynoed_eval_fixed_controls(self._h, _ptr(basis), _ptr(designs), n, _ptr(crit),
                                                  _ptr(grad), _ptr(fim)), self._h, self._L)

    def kernel_stats(self, with_grad=True) -> dict:
        ks = KernelStats()
        _check(self._L.dynoed_kernel_stats_query(self._h, int(with_grad), C.byref(ks)), self._h, self._L)   # 2: NO_CONTROLS kernel
        return {k: getattr(ks, k) for k, _ in KernelStats._fields_}

    # convenience: host numpy in, numpy out
    def evaluate(self, designs, grad=True, fim=True, flags=0):
        designs = np.ascontiguousarray(np.atleast_2d(designs), dtype=np.float64)
        n = designs.shape[0]
        crit = np.empty(n)
        g = np.empty((n, self.n_var)) if grad else None
        f = np.empty((n, self.info.nF)) if fim else None
        self.eval_batch(designs, crit, g, f, flags=flags)
        return crit, g, f


def criterion_batch(criterion: int, F_packed, tau=None, with_gradient=False, device: int = 0):
    """c(Symmetric(F + tau I)) for a batch of packed symmetric matrices — on the GPU."""
    F_packed = np.ascontiguousarray(np.atleast_2d(F_packed), dtype=np.float64)
    b, nF = F_packed.shape
    n = int(round((np.sqrt(8 * nF + 1) - 1) / 2))
    assert n * (n + 1) // 2 == nF
    t = None if tau is None else np.ascontiguousarray(np.broadcast_to(np.asarray(tau, dtype=np.float64), (b,)))
    val = np.empty(b)
    dL = np.empty((b, nF)) if with_gradient else None
    _check(lib().dynoed_criterion_batch(device, int(criterion), n, _ptr(F_packed), _ptr(t), b,
                                        _ptr(val), _ptr(dL)))
    return (val, dL) if with_gradient else val


def gpu_numa_cpus(device: int = 0):
    """(NUMA node, sorted CPU list) of the host CPUs closest to CUDA device `device`, from sysfs; (None, [])
    when the box does not expose it.  A host thread that feeds the GPU from page-locked memory should run,
    and allocate that memory, on these CPUs: `os.sched_setaffinity(0, cpus)` before the first allocation."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(device), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None, []
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        return node, cpus
    except Exception:
        return None, []


def fp64_peak(device: int = 0, stream_ptr=None):
    t, ms = C.c_double(), C.c_double()
    _check(lib().dynoed_fp64_peak(device, C.c_void_p(stream_ptr) if stream_ptr else None,
                                  C.byref(t), C.byref(ms)))
    return float(t.value), float(ms.value)


def fp64_microbench(variant: int, device: int = 0) -> float:
    """Thread-level FP64 instructions per second (in 1e9) for an operand pattern (see dynoed.h)."""
    g = C.c_double()
    _check(lib().dynoed_fp64_microbench(device, int(variant), C.byref(g)))
    return float(g.value)
