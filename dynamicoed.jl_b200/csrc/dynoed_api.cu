// dynoed_api.cu — the C ABI of libdynoed_cuda.so (see include/dynoed.h) and the host runtime
// around the kernel template: model registry, per-problem device tables, scratch ownership,
// launch configuration, the chunked H2D / compute / D2H pipeline for host buffers, arg-min.
//
// There is deliberately no CPU code path in here: every compute entry point launches CUDA
// kernels or fails with DYNOED_ECUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/dynoed.h"
#include "dynoed_kernels.cuh"
#include "dynoed_rosenbrock.cuh"
#ifdef DYNOED_MODELS_HEADER
#include DYNOED_MODELS_HEADER
#else
#include "generated/models.cuh"
#endif

namespace dynoed {

static thread_local std::string g_last_error;

// method + kFwdOnly selects the forward-only gradient kernel of an explicit tableau (DYNOED_GRAD_NO_CONTROLS)
constexpr int kFwdOnly = 16;
// method | kCtrlGrad: Rosenbrock gradient kernel that also differentiates with respect to control values
constexpr int kCtrlGrad = 32;
// method | kDirPar: direction-parallel gradient kernel for small batches (dir_eval_kernel)
constexpr int kDirPar = 64;
// method | kIc2: Rosenbrock gradient with second-order sensitivities for unknown initial conditions (ros_ic_kernel)
constexpr int kIc2 = 128;
// method | kTimePar: time-parallel gradient kernel for the smallest batches (tp_eval_kernel, one CTA per design)
constexpr int kTimePar = 256;
// ... | kTpSpill: with its per-step tables in global scratch instead of shared memory
constexpr int kTpSpill = 512;
constexpr int kDirBlock = 32;
// host calls of at most this many designs let the kernel read / write page-locked host memory directly
constexpr int64_t kSmallZeroCopy = 16;

struct ModelEntry {
    const char* name;
    dynoed_model_info info;
    cudaError_t (*launch)(int method, bool grad, bool ucoef, bool pdl, const EvalParams& P, int grid, int block,
                          size_t smem, cudaStream_t st);
    cudaError_t (*attrs)(int method, bool grad, bool ucoef, int block, size_t smem, cudaFuncAttributes* fa,
                         int* blocks_per_sm);
    cudaError_t (*info_gain)(const EvalParams& P, const double* X, double* Pout, double* Piout, cudaStream_t st);
    int flops_ros_jac, flops_ros_ling, flops_ros_linp, flops_ros_hvp;   // emitted Rosenbrock pieces (FMA = 2)
    int nh, flops_ic2_col, flops_ic2_q;   // unique second-order columns and their emitted pieces (all columns, one stage)
    bool ic2_ok;
    int ic2_gcol[kMaxIC];                 // G column that is d x / d ic_a
    int tp_nd;                            // tp_eval_kernel: adjoint lanes (nz + n_ctrl), 0 = kernel not built
    size_t (*tp_smem)(int n_var, int n_intervals, int n_steps, int n_ctrl_values, bool spill);
    size_t (*tp_spill_bytes)(int n_steps);   // global scratch per design in spill mode
};

template <class K>
static cudaError_t launch_kernel(K kern, bool pdl, const EvalParams& P, int grid, int block, size_t smem,
                                 cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, P);
}

template <class M, class Tab, bool GRAD, bool UC>
static cudaError_t launch_one(bool pdl, const EvalParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    return launch_kernel(oed_eval_kernel<M, Tab, GRAD, UC>, pdl, P, grid, block, smem, st);
}

template <class M, class Tab, bool GRAD, bool UC>
static cudaError_t attrs_one(int block, size_t smem, cudaFuncAttributes* fa, int* bps) {
    auto kern = oed_eval_kernel<M, Tab, GRAD, UC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(fa, kern);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, block, smem);
}

template <class M, class Tab>
static cudaError_t launch_tab(bool grad, bool uc, bool pdl, const EvalParams& P, int grid, int block, size_t smem,
                              cudaStream_t st) {
    if (grad) return uc ? launch_one<M, Tab, true, true>(pdl, P, grid, block, smem, st)
                        : launch_one<M, Tab, true, false>(pdl, P, grid, block, smem, st);
    return uc ? launch_one<M, Tab, false, true>(pdl, P, grid, block, smem, st)
              : launch_one<M, Tab, false, false>(pdl, P, grid, block, smem, st);
}

// directions per integration pass of the forward-mode gradient kernels (DYNOED_ROS_NDIR at build time)
#ifndef DYNOED_ROS_NDIR
#define DYNOED_ROS_NDIR 2
#endif
// Up to DYNOED_ROS_NDIR directions share one pass (and its value computations); four-stage schemes keep
// one direction per pass — with two the stage vectors of models of chem6's size spill so much that the
// pass is slower than two single-direction passes (measured on B200: RODAS3 2.4e6 vs 2.8e6 designs/s).
template <class M, bool CG, class Stepper>
constexpr int ros_ndir() {
    constexpr int cap = Stepper::STAGES >= 4 ? 1 : DYNOED_ROS_NDIR;
    return (CG && M::NCTRL > 0) ? cap : (M::NIC > 0 && M::NIC < cap ? M::NIC : M::NIC > 0 ? cap : 1);
}

// CG: differentiate with respect to control values too (Rosenbrock steppers)
template <class M, class Stepper, bool GRAD, bool CG = false>
static cudaError_t launch_ros(bool pdl, const EvalParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    return launch_kernel(ros_eval_kernel<M, Stepper, GRAD, CG, ros_ndir<M, CG, Stepper>()>, pdl, P, grid, block, smem, st);
}

template <class M, class Stepper, bool GRAD, bool CG = false>
static cudaError_t attrs_ros(int block, size_t smem, cudaFuncAttributes* fa, int* bps) {
    auto kern = ros_eval_kernel<M, Stepper, GRAD, CG, ros_ndir<M, CG, Stepper>()>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(fa, kern);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, block, smem);
}

// forward-only gradient mode of an explicit tableau (DYNOED_GRAD_NO_CONTROLS): models without unknown initial
// conditions run the evaluation kernel's own forward sweep with the quadratures kept (oed_wgrad_kernel); the
// others the generic forward-only kernel, which carries dual parts for the initial conditions
template <class M, class Tab>
static cudaError_t launch_fwdonly(bool uc, bool pdl, const EvalParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    if constexpr (M::NIC == 0)
        return uc ? launch_kernel(oed_wgrad_kernel<M, Tab, true>, pdl, P, grid, block, smem, st)
                  : launch_kernel(oed_wgrad_kernel<M, Tab, false>, pdl, P, grid, block, smem, st);
    else
        return launch_ros<M, ExplicitStepper<Tab>, true>(pdl, P, grid, block, smem, st);
}
template <class M, class Tab, bool UC>
static cudaError_t attrs_wgrad(int block, size_t smem, cudaFuncAttributes* fa, int* bps) {
    auto kern = oed_wgrad_kernel<M, Tab, UC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(fa, kern);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, block, smem);
}
template <class M, class Tab>
static cudaError_t attrs_fwdonly(bool uc, int block, size_t smem, cudaFuncAttributes* fa, int* bps) {
    if constexpr (M::NIC == 0)
        return uc ? attrs_wgrad<M, Tab, true>(block, smem, fa, bps) : attrs_wgrad<M, Tab, false>(block, smem, fa, bps);
    else
        return attrs_ros<M, ExplicitStepper<Tab>, true>(block, smem, fa, bps);
}

// Rosenbrock gradient of a model with unknown initial conditions and no controls: ros_ic_kernel
template <class M, class RT>
static cudaError_t launch_ic(bool pdl, const EvalParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    if constexpr (M::IC2_OK) return launch_kernel(ros_ic_kernel<M, RT>, pdl, P, grid, block, smem, st);
    else return cudaErrorInvalidDeviceFunction;
}
template <class M, class RT>
static cudaError_t attrs_ic(int block, size_t smem, cudaFuncAttributes* fa, int* bps) {
    if constexpr (M::IC2_OK) {
        auto kern = ros_ic_kernel<M, RT>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncGetAttributes(fa, kern);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, block, smem);
    } else {
        return cudaErrorInvalidDeviceFunction;
    }
}

template <class M, class Stepper>
static cudaError_t launch_dir(const EvalParams& P0, int grid, cudaStream_t st) {
    // the design rows of one CTA's lanes are staged in shared memory when they fit the default 48 KB
    EvalParams P = P0;
    const int nl = 1 + M::NIC + P.n_ctrl_values;
    const size_t rows = (size_t)(kDirBlock + nl - 1) / nl + 1;
    size_t smem = rows * (P.n_var + (size_t)P.n_intervals * M::NOBS * M::NF) * sizeof(double) +
                  sizeof(int) * ((size_t)(M::NCTRL + M::NOBS) * P.n_intervals + P.n_intervals + 1);
    P.tile = smem <= 48 * 1024 ? (int)rows : 0;
    if (P.tile == 0) smem = 0;
    dir_eval_kernel<M, Stepper><<<grid, kDirBlock, smem, st>>>(P);
    return cudaGetLastError();
}

// only models that have something besides w and tau to differentiate ever run the kernel (see dir_ok):
// the others do not instantiate it
template <class M>
constexpr bool tp_built() { return M::NZ + 1 <= 32 && M::NIC + M::NCTRL > 0; }

template <class M>
static size_t tp_smem_model(int n_var, int N, int n_steps, int ncv, bool spill) { return tp_smem_bytes<M>(n_var, N, n_steps, ncv, spill); }
template <class M>
static size_t tp_spill_model(int n_steps) { return sizeof(double) * tp_spill_doubles<M>(n_steps); }

template <class M, class Stepper, bool SPILL>
static cudaError_t launch_tp_one(const EvalParams& P, int grid, cudaStream_t st) {
    const size_t smem = tp_smem_bytes<M>(P.n_var, P.n_intervals, P.n_steps, P.n_ctrl_values, SPILL);
    static size_t smem_set[64] = {0};      // opt-in dynamic shared memory, raised once per device and size
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > smem_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(tp_eval_kernel<M, Stepper, SPILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set[dev & 63] = smem;
    }
    tp_eval_kernel<M, Stepper, SPILL><<<grid, kTpBlock, smem, st>>>(P);
    return cudaGetLastError();
}
template <class M, class Stepper>
static cudaError_t launch_tp(bool spill, const EvalParams& P, int grid, cudaStream_t st) {
    if constexpr (tp_built<M>()) {
        return spill ? launch_tp_one<M, Stepper, true>(P, grid, st) : launch_tp_one<M, Stepper, false>(P, grid, st);
    } else {
        return cudaErrorInvalidDeviceFunction;
    }
}

template <class M>
static cudaError_t launch_model(int method, bool grad, bool uc, bool pdl, const EvalParams& P, int grid, int block,
                                size_t smem, cudaStream_t st) {
    if (method & kTimePar) {
        const bool sp = (method & kTpSpill) != 0;
        switch (method & 15) {
            case DYNOED_RK4: return launch_tp<M, ExplicitStepper<TabRK4>>(sp, P, grid, st);
            case DYNOED_TSIT5: return launch_tp<M, ExplicitStepper<TabTsit5>>(sp, P, grid, st);
            case DYNOED_ROS2: return launch_tp<M, RosenbrockStepper<TabROS2>>(sp, P, grid, st);
            case DYNOED_ROS3P: return launch_tp<M, RosenbrockStepper<TabROS3P>>(sp, P, grid, st);
            default: return launch_tp<M, RosenbrockStepper<TabRODAS3>>(sp, P, grid, st);
        }
    }
    if (method & kDirPar) {
        switch (method & 15) {
            case DYNOED_RK4: return launch_dir<M, ExplicitStepper<TabRK4>>(P, grid, st);
            case DYNOED_TSIT5: return launch_dir<M, ExplicitStepper<TabTsit5>>(P, grid, st);
            case DYNOED_ROS2: return launch_dir<M, RosenbrockStepper<TabROS2>>(P, grid, st);
            case DYNOED_ROS3P: return launch_dir<M, RosenbrockStepper<TabROS3P>>(P, grid, st);
            default: return launch_dir<M, RosenbrockStepper<TabRODAS3>>(P, grid, st);
        }
    }
    if (method & kIc2) {
        switch (method & 15) {
            case DYNOED_ROS2: return launch_ic<M, TabROS2>(pdl, P, grid, block, smem, st);
            case DYNOED_ROS3P: return launch_ic<M, TabROS3P>(pdl, P, grid, block, smem, st);
            default: return launch_ic<M, TabRODAS3>(pdl, P, grid, block, smem, st);
        }
    }
    const bool cg = (method & kCtrlGrad) != 0;
    method &= ~kCtrlGrad;
    if (method == DYNOED_RK4) return launch_tab<M, TabRK4>(grad, uc, pdl, P, grid, block, smem, st);
    if (method == DYNOED_TSIT5) return launch_tab<M, TabTsit5>(grad, uc, pdl, P, grid, block, smem, st);
    if (method == DYNOED_ROS2)
        return !grad ? launch_ros<M, RosenbrockStepper<TabROS2>, false>(pdl, P, grid, block, smem, st)
               : cg ? launch_ros<M, RosenbrockStepper<TabROS2>, true, (M::NCTRL > 0)>(pdl, P, grid, block, smem, st)
                    : launch_ros<M, RosenbrockStepper<TabROS2>, true>(pdl, P, grid, block, smem, st);
    if (method == DYNOED_ROS3P)
        return !grad ? launch_ros<M, RosenbrockStepper<TabROS3P>, false>(pdl, P, grid, block, smem, st)
               : cg ? launch_ros<M, RosenbrockStepper<TabROS3P>, true, (M::NCTRL > 0)>(pdl, P, grid, block, smem, st)
                    : launch_ros<M, RosenbrockStepper<TabROS3P>, true>(pdl, P, grid, block, smem, st);
    if (method == DYNOED_RODAS3)
        return !grad ? launch_ros<M, RosenbrockStepper<TabRODAS3>, false>(pdl, P, grid, block, smem, st)
               : cg ? launch_ros<M, RosenbrockStepper<TabRODAS3>, true, (M::NCTRL > 0)>(pdl, P, grid, block, smem, st)
                    : launch_ros<M, RosenbrockStepper<TabRODAS3>, true>(pdl, P, grid, block, smem, st);
    // forward-only gradient mode of the explicit tableaus (kFwdOnly + method)
    if (method == kFwdOnly + DYNOED_RK4) return launch_fwdonly<M, TabRK4>(uc, pdl, P, grid, block, smem, st);
    return launch_fwdonly<M, TabTsit5>(uc, pdl, P, grid, block, smem, st);
}

template <class M, class Tab>
static cudaError_t attrs_tab(bool grad, bool uc, int block, size_t smem, cudaFuncAttributes* fa, int* bps) {
    if (grad) return uc ? attrs_one<M, Tab, true, true>(block, smem, fa, bps)
                        : attrs_one<M, Tab, true, false>(block, smem, fa, bps);
    return uc ? attrs_one<M, Tab, false, true>(block, smem, fa, bps)
              : attrs_one<M, Tab, false, false>(block, smem, fa, bps);
}

template <class M>
static cudaError_t attrs_model(int method, bool grad, bool uc, int block, size_t smem, cudaFuncAttributes* fa,
                               int* bps) {
    if (method & kIc2) {
        switch (method & 15) {
            case DYNOED_ROS2: return attrs_ic<M, TabROS2>(block, smem, fa, bps);
            case DYNOED_ROS3P: return attrs_ic<M, TabROS3P>(block, smem, fa, bps);
            default: return attrs_ic<M, TabRODAS3>(block, smem, fa, bps);
        }
    }
    const bool cg = (method & kCtrlGrad) != 0;
    method &= ~kCtrlGrad;
    if (method == DYNOED_RK4) return attrs_tab<M, TabRK4>(grad, uc, block, smem, fa, bps);
    if (method == DYNOED_TSIT5) return attrs_tab<M, TabTsit5>(grad, uc, block, smem, fa, bps);
    if (method == DYNOED_ROS2)
        return !grad ? attrs_ros<M, RosenbrockStepper<TabROS2>, false>(block, smem, fa, bps)
               : cg ? attrs_ros<M, RosenbrockStepper<TabROS2>, true, (M::NCTRL > 0)>(block, smem, fa, bps)
                    : attrs_ros<M, RosenbrockStepper<TabROS2>, true>(block, smem, fa, bps);
    if (method == DYNOED_ROS3P)
        return !grad ? attrs_ros<M, RosenbrockStepper<TabROS3P>, false>(block, smem, fa, bps)
               : cg ? attrs_ros<M, RosenbrockStepper<TabROS3P>, true, (M::NCTRL > 0)>(block, smem, fa, bps)
                    : attrs_ros<M, RosenbrockStepper<TabROS3P>, true>(block, smem, fa, bps);
    if (method == DYNOED_RODAS3)
        return !grad ? attrs_ros<M, RosenbrockStepper<TabRODAS3>, false>(block, smem, fa, bps)
               : cg ? attrs_ros<M, RosenbrockStepper<TabRODAS3>, true, (M::NCTRL > 0)>(block, smem, fa, bps)
                    : attrs_ros<M, RosenbrockStepper<TabRODAS3>, true>(block, smem, fa, bps);
    if (method == kFwdOnly + DYNOED_RK4) return attrs_fwdonly<M, TabRK4>(uc, block, smem, fa, bps);
    return attrs_fwdonly<M, TabTsit5>(uc, block, smem, fa, bps);
}

template <class M>
static cudaError_t info_gain_model(const EvalParams& P, const double* X, double* Pout, double* Piout,
                                   cudaStream_t st) {
    const long long total = P.n * (P.n_intervals + 1);
    info_gain_kernel<M><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(P, X, Pout, Piout);
    return cudaGetLastError();
}

template <class M>
static ModelEntry make_entry() {
    static_assert(M::NX <= kMaxX && M::NX * M::NP <= kMaxG && M::NTH <= kMaxTheta, "model too large");
    static_assert(M::NCTRL <= kMaxCtrl && M::NOBS <= kMaxObs && M::NIC <= kMaxIC, "model too large");
    ModelEntry e{};
    e.name = M::NAME;
    dynoed_model_info& i = e.info;
    memset(&i, 0, sizeof(i));
    i.nx = M::NX; i.np = M::NP; i.n_obs = M::NOBS; i.n_ctrl = M::NCTRL; i.n_ic = M::NIC;
    i.n_theta = M::NTH; i.nz = M::NZ; i.nF = M::NF;
    for (int k = 0; k < M::NIC; ++k) i.ic_state[k] = M::ic_state(k);
    i.autonomous = M::AUTONOMOUS ? 1 : 0;
    i.flops_primal_z = M::FLOPS_PRIMAL_Z;
    i.flops_adjoint = M::FLOPS_ADJOINT; i.flops_consts = M::FLOPS_CONSTS; i.flops_ros_f = M::FLOPS_ROS_F;
    i.flops_primal = M::FLOPS_PRIMAL_Z + M::FLOPS_OBS;
    for (int k = 0; k < M::NX; ++k) i.x0[k] = M::default_x0()[k];
    for (int k = 0; k < M::NX * M::NP; ++k) i.G0[k] = M::default_G0()[k];
    for (int k = 0; k < M::NTH; ++k) i.theta[k] = M::default_theta()[k];
    e.flops_ros_jac = M::FLOPS_ROS_JAC; e.flops_ros_ling = M::FLOPS_ROS_LING;
    e.flops_ros_linp = M::FLOPS_ROS_LINP; e.flops_ros_hvp = M::FLOPS_ROS_HVP;
    e.nh = M::NH; e.flops_ic2_col = M::FLOPS_IC2_COL; e.flops_ic2_q = M::FLOPS_IC2_Q; e.ic2_ok = M::IC2_OK;
    for (int k = 0; k < M::NIC; ++k) e.ic2_gcol[k] = M::ic2_gcol(k);

    e.tp_nd = tp_built<M>() ? M::NZ + M::NCTRL : 0;
    e.tp_smem = &tp_smem_model<M>;
    e.tp_spill_bytes = &tp_spill_model<M>;
    e.launch = &launch_model<M>;
    e.attrs = &attrs_model<M>;
    e.info_gain = &info_gain_model<M>;
    return e;
}

static const std::vector<ModelEntry>& registry() {
    static const std::vector<ModelEntry> r = [] {
        std::vector<ModelEntry> v;
#define DYNOED_REG(M) v.push_back(make_entry<M>());
        DYNOED_FOR_EACH_MODEL(DYNOED_REG)
#undef DYNOED_REG
        return v;
    }();
    return r;
}

static const ModelEntry* find_model(const char* name) {
    if (!name) return nullptr;
    for (const auto& e : registry())
        if (strcmp(e.name, name) == 0) return &e;
    return nullptr;
}

// Serial number of the last evaluation launch this library put on a stream (any context).  A
// launch may only overlap its predecessor if that predecessor is the SAME context's previous
// launch (whose buffers it was checked against); foreign kernels in between are harmless because
// they do not trigger programmatic completion early.
static std::mutex g_stream_mu;
static std::vector<std::pair<cudaStream_t, uint64_t>> g_stream_serial;
static uint64_t g_serial = 0;
static uint64_t stream_last_serial(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    for (auto& e : g_stream_serial)
        if (e.first == st) return e.second;
    return 0;
}
static uint64_t stream_new_serial(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    const uint64_t v = ++g_serial;
    for (auto& e : g_stream_serial)
        if (e.first == st) { e.second = v; return v; }
    if (g_stream_serial.size() > 256) g_stream_serial.clear();
    g_stream_serial.emplace_back(st, v);
    return v;
}

constexpr int kSlots = 4;   // pipeline depth of the host-buffer path
constexpr int kRing = 64;   // arg-min records kept for deferred (batched) multi-GPU gathers

struct Slot {
    cudaStream_t stream = nullptr;
    double *d_designs = nullptr, *d_crit = nullptr, *d_grad = nullptr, *d_fim = nullptr;
    double* traj = nullptr;
    size_t traj_bytes = 0;
    int64_t cap = 0;   // designs
    // page-locked staging for PAGEABLE caller arrays (ordinary malloc / numpy / Julia arrays)
    double *h_designs = nullptr, *h_crit = nullptr, *h_grad = nullptr, *h_fim = nullptr;
    double *hd_designs = nullptr, *hd_crit = nullptr, *hd_grad = nullptr, *hd_fim = nullptr;   // their device addresses
    int64_t hcap = 0;
    cudaEvent_t done = nullptr;
};

// memcpy split over a few host threads: one core moves ~10 GB/s, a PCIe Gen5 link wants ~50
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    if (bytes < (2u << 20)) { memcpy(dst, src, bytes); return; }
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nt = (int)std::min<size_t>(std::min<unsigned>(8u, std::max(1u, hw / 2)), bytes / (1u << 20));
    if (nt <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 63) & ~(size_t)63;
    for (int t = 0; t < nt; ++t) {
        const size_t lo = std::min(bytes, (size_t)t * per), hi = std::min(bytes, lo + per);
        if (hi > lo) th.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    for (auto& x : th) x.join();
}

}  // namespace dynoed

using namespace dynoed;

struct dynoed_ctx {
    int device = 0;
    const ModelEntry* model = nullptr;
    int method = DYNOED_RK4;
    int stages = 4;
    int criterion = 0;
    int n_intervals = 0, n_steps = 0, n_var = 0, tile = 128;
    int sm_count = 0;
    size_t smem_optin = 0, smem_per_sm = 0;
    int w_len[kMaxObs] = {0};   // length of each measurement-weight block of the design vector
    bool ucoef = false;    // step coefficients served from the constant bank (uniform registers)
    bool rosenbrock = false;
    bool ic2 = false;      // gradient kernel = ros_ic_kernel (Rosenbrock, unknown initial conditions, no controls)
    int bps[3] = {0, 0, 0};   // blocks per SM: objective only / gradient / forward-only gradient (explicit RK)
    size_t smem_floor = 0; // dynamic shared memory padding that caps the resident CTAs per SM
    cudaFuncAttributes fa[3];
    EvalParams base;       // tables + layout + model data (batch pointers filled per launch)
    int* d_first_step = nullptr;
    double* d_step_t = nullptr;
    double* d_step_h = nullptr;
    double* d_step_ih = nullptr;
    int* d_indicators = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // device-pointer path scratch, double-buffered by launch parity: consecutive device-path
    // launches may overlap (programmatic dependent launch), so launch k and k+1 never share
    // checkpoint scratch, arg-min records, ticket or result slot
    double* traj[2] = {nullptr, nullptr};
    size_t traj_bytes[2] = {0, 0};
    double* part_val[2] = {nullptr, nullptr};
    long long* part_idx[2] = {nullptr, nullptr};
    int part_cap[2] = {0, 0};
    long long* d_best = nullptr;       // [2 parities][2]: {criterion value (double bits), design index}
    long long* d_ring = nullptr;       // [kRing][2]: the same record for the last kRing batches
    int64_t batch_seq = 0;             // batches evaluated so far (ring slot = seq % kRing)
    unsigned int* d_ticket = nullptr;  // [2]
    unsigned long long* d_guard = nullptr;   // [2][4]: calls finished with the set, waits, timed-out waits (scratch_acquire)
    unsigned long long set_calls[2] = {0, 0};   // calls handed each scratch set so far (incl. the current one)
    unsigned long long call_serial = 0;      // serial of the most recent evaluation call (> 0)
    uint32_t call_flags = 0;                 // flags of the call being served
    std::vector<int> h_first;                // host copy of first_step[N+1]
    size_t l2_bytes = 0;
    double l2_hot_frac = 1.0;                // share of L2 the top of the checkpoint stack may occupy
    cudaStream_t aux_stream = nullptr;       // selection helpers run here when set (dynoed_set_aux_stream)
    cudaEvent_t ev_aux = nullptr;
    struct HessScratch {               // dynoed_hessian_batch: perturbed batch + staging, grown on demand
        double *pert = nullptr, *pcrit = nullptr, *pgrad = nullptr, *x = nullptr, *c = nullptr, *g = nullptr, *H = nullptr;
        int64_t cap = 0;               // designs per chunk the buffers hold
    } hs;
    bool dir_ok = false;               // direction-parallel kernel usable for small gradient batches
    int dir_nl = 1;                    // its lanes per design: 1 + unknown ICs + control values
    int64_t dir_max_lanes = 0;
    bool tp_ok = false;                // time-parallel kernel usable for the smallest gradient batches
    bool tp_spill = false;             // ... with its per-step tables in global scratch (they exceed shared memory)
    int64_t tp_max_n = 0;
    int64_t call_n = 0;                // designs of the call being served
    bool call_xgrid = false;           // ... and it writes the trajectory (objective-only kernels of the large-batch path)
    int kidx_req = 1;                  // gradient kernel requested by the current call (1 adjoint/Rosenbrock, 2 forward-only)
    int parity = 0;                    // scratch set of the NEXT device-path launch
    int last_parity = 0;               // scratch set that holds the most recent batch's arg-min
    struct Prev {                      // the previous device-path launch (overlap eligibility)
        bool valid = false, full = false;
        int kidx = 0;
        uint64_t serial = 0;
        cudaStream_t stream = nullptr;
        const char *in0 = nullptr, *in1 = nullptr;
        const char* out0[3] = {nullptr, nullptr, nullptr};
        const char* out1[3] = {nullptr, nullptr, nullptr};
    } prev;
    // pipelined host calls (dynoed_eval_batch_host_async): completion event per scratch parity
    cudaEvent_t call_done[2] = {nullptr, nullptr};
    int64_t call_of_parity[2] = {-1, -1};   // id of the last host call that used each parity
    int64_t host_seq = 0;                   // host calls enqueued so far (ids are 1-based)
    bool stream_dirty = true;               // work other than host calls was put on ctx->stream since the last host call
    bool pdl_enabled = true;
    bool zero_copy = true;             // store crit / fim straight into page-locked host arrays
    bool host_staging = true;          // pageable caller arrays go through page-locked staging
    int host_chunks = 8;               // equal chunks; 0 (DYNOED_HOST_CHUNKS=0) selects the ramped schedule
    bool have_batch = false;
    Slot slots[kSlots];
    int64_t launches = 0;
    int64_t pdl_launches = 0;
    cudaEvent_t ev_order = nullptr;    // host path: orders the slot streams after / before the ctx stream
    int64_t dir_launches = 0;
    int64_t tp_launches = 0;
    size_t scratch_bytes = 0;
    // optional per-launch timing of the evaluation kernel (device-pointer path)
    bool profiling = false;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    double kernel_ms_sum = 0.0;
    int64_t kernel_ms_count = 0;
    bool ev_pending = false;
    std::string err;
};

static int fail(dynoed_ctx* ctx, int code, const std::string& msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, DYNOED_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

extern "C" {

int dynoed_abi_version(void) { return DYNOED_ABI_VERSION; }
int dynoed_model_count(void) { return (int)registry().size(); }
const char* dynoed_model_name(int i) {
    if (i < 0 || i >= (int)registry().size()) return nullptr;
    return registry()[i].name;
}
int dynoed_model_query(const char* name, dynoed_model_info* out) {
    const ModelEntry* e = find_model(name);
    if (!e) return fail(nullptr, DYNOED_ENOMODEL, std::string("unknown model: ") + (name ? name : "(null)"));
    if (!out) return fail(nullptr, DYNOED_EINVAL, "out is NULL");
    *out = e->info;
    return DYNOED_OK;
}

const char* dynoed_last_error(const dynoed_ctx* ctx) {
    if (ctx && !ctx->err.empty()) return ctx->err.c_str();
    return g_last_error.c_str();
}

static int ensure_partials(dynoed_ctx* ctx, int need, int par) {
    if (need <= ctx->part_cap[par]) return DYNOED_OK;
    CU(cudaDeviceSynchronize());
    if (ctx->part_val[par]) cudaFree(ctx->part_val[par]);
    if (ctx->part_idx[par]) cudaFree(ctx->part_idx[par]);
    ctx->part_val[par] = nullptr; ctx->part_idx[par] = nullptr;
    int cap = std::max(need, 4096);
    CU(cudaMalloc(&ctx->part_val[par], sizeof(double) * cap));
    CU(cudaMalloc(&ctx->part_idx[par], sizeof(long long) * cap));
    ctx->part_cap[par] = cap;
    return DYNOED_OK;
}

// per-CTA scratch of a gradient launch: (x, G) checkpoints of every step (explicit RK adjoint) or
// the per-interval unweighted quadratures (Rosenbrock)
static size_t traj_bytes_for(const dynoed_ctx* ctx, int grid, int kidx = 1) {
    const auto& mi = ctx->model->info;
    const size_t per_lane = (ctx->rosenbrock || kidx == 2)
                                ? (size_t)ctx->n_intervals * mi.n_obs * mi.nF + ctx->base.n_ctrl_values
                                : (size_t)ctx->n_steps * mi.nz + (size_t)ctx->n_intervals * mi.n_obs * mi.nF;
    return (size_t)grid * per_lane * kMaxTile * sizeof(double);
}

// 0: objective only, 1: gradient (adjoint sweep, or the Rosenbrock kernel), 2: forward-only gradient
static int kernel_index(const dynoed_ctx* ctx, bool grad, uint32_t flags) {
    (void)ctx;
    if (!grad) return 0;
    return (flags & DYNOED_GRAD_NO_CONTROLS) ? 2 : 1;
}

// the method code the model's launch / attrs dispatch understands, per kernel index
static int method_code(const dynoed_ctx* ctx, int kidx) {
    if (ctx->rosenbrock && kidx == 1 && ctx->ic2) return ctx->method | kIc2;
    if (ctx->rosenbrock) return (kidx == 1 && ctx->model->info.n_ctrl > 0) ? (ctx->method | kCtrlGrad) : ctx->method;
    return kidx == 2 ? kFwdOnly + ctx->method : ctx->method;
}

// kidx 1 of an ic2 context: behind the design tile every lane owns Ic2Layout::SLOTS doubles (stage
// increments of (x, G) and the second-order sensitivity columns), strided by kMaxTile
static size_t smem_bytes(const dynoed_ctx* ctx, int kidx = 0) {
    size_t need = 128 + (size_t)ctx->tile * ctx->n_var * sizeof(double);
    if (kidx == 1 && ctx->ic2) {
        const auto& mi = ctx->model->info;
        const size_t slots = (size_t)ctx->stages * mi.nz + (size_t)ctx->model->nh * mi.nx;
        const size_t stride = ctx->stages >= 4 ? 64 : kMaxTile;      // ic_stride<RT>()
        need = 128 + ((((size_t)ctx->tile * ctx->n_var + 15) & ~(size_t)15) + slots * stride) * sizeof(double);
    }
    return std::max(ctx->smem_floor, need);
}

// Small gradient batches (the optimiser callback, finite-difference Hessians) are bound by the
// sequential depth of one design, not by arithmetic: they run on the direction-parallel kernel
// (1 + n_dir lanes per design, forward sweep only) while its lanes stay below three warps per SM.
static bool dir_eligible(const dynoed_ctx* ctx, int64_t n, int kidx) {
    // decided on the size of the whole CALL (a chunk of a large host batch must not switch kernels)
    return kidx == 1 && ctx->dir_ok && n > 0 && std::max(n, ctx->call_n) * ctx->dir_nl <= ctx->dir_max_lanes;
}

// The smallest of them (one optimiser callback, the 2 n_var + 1 designs of a finite-difference Hessian)
// run on the time-parallel kernel: one CTA per design, see tp_eval_kernel.
static bool tp_eligible(const dynoed_ctx* ctx, int64_t n, int kidx) {
    return ctx->tp_ok && dir_eligible(ctx, n, kidx) && std::max(n, ctx->call_n) <= ctx->tp_max_n;
}

// ... and the smallest OBJECTIVE-ONLY calls too (the line-search evaluations of an optimiser): same kernel
// without its dual / adjoint phases, reading and writing page-locked host memory directly for host calls
// instead of three small DMA transfers.  Not with spilled tables (no scratch is sized for objective calls).
static bool tp_obj_eligible(const dynoed_ctx* ctx, int64_t n, int kidx) {
    return kidx == 0 && ctx->tp_ok && !ctx->tp_spill && !ctx->call_xgrid && n > 0 &&
           std::max(n, ctx->call_n) <= ctx->tp_max_n;
}

static int grid_for(const dynoed_ctx* ctx, int64_t n, int kidx) {
    if (tp_eligible(ctx, n, kidx) || tp_obj_eligible(ctx, n, kidx)) return (int)n;
    if (dir_eligible(ctx, n, kidx)) return (int)((n * ctx->dir_nl + kDirBlock - 1) / kDirBlock);
    const int64_t ntiles = (n + ctx->tile - 1) / ctx->tile;
    const int64_t resident = (int64_t)ctx->sm_count * std::max(ctx->bps[kidx], 1);
    return (int)std::max<int64_t>(1, std::min(ntiles, resident));
}

// scratch of one gradient launch of n designs
static size_t traj_need(const dynoed_ctx* ctx, int64_t n, int kidx) {
    if (tp_eligible(ctx, n, kidx))                // everything lives in shared memory unless the tables spill
        return ctx->tp_spill ? (size_t)n * ctx->model->tp_spill_bytes(ctx->n_steps) : 0;
    if (dir_eligible(ctx, n, kidx))
        return sizeof(double) * (size_t)n * ctx->n_intervals * ctx->model->info.n_obs * ctx->model->info.nF;
    return traj_bytes_for(ctx, grid_for(ctx, n, kidx), kidx);
}


int dynoed_create(const dynoed_problem_desc* d, int device, dynoed_ctx** out) {
    dynoed_ctx* ctx = nullptr;
    if (!d || !out) return fail(nullptr, DYNOED_EINVAL, "desc/out is NULL");
    if (d->struct_size != sizeof(dynoed_problem_desc))
        return fail(nullptr, DYNOED_EINVAL, "dynoed_problem_desc.struct_size mismatch (ABI)");
    const ModelEntry* m = find_model(d->model);
    if (!m) return fail(nullptr, DYNOED_ENOMODEL, std::string("unknown model: ") + (d->model ? d->model : "(null)"));
    const dynoed_model_info& mi = m->info;
    if (d->criterion < 0 || d->criterion >= kNumCriteria) return fail(nullptr, DYNOED_EINVAL, "criterion out of range");
    if (d->method < DYNOED_RK4 || d->method > DYNOED_RODAS3)
        return fail(nullptr, DYNOED_EUNSUPPORTED, "unsupported integration method");
    const bool rosenbrock = d->method == DYNOED_ROS2 || d->method == DYNOED_ROS3P || d->method == DYNOED_RODAS3;
    if (rosenbrock && !mi.autonomous)
        return fail(nullptr, DYNOED_EUNSUPPORTED, "the Rosenbrock integrators need an autonomous right-hand side");
    const int N = d->n_intervals;
    if (N <= 0 || !d->tgrid || !d->indicators) return fail(nullptr, DYNOED_EINVAL, "empty grid");
    if ((mi.n_ctrl && !d->ctrl_offsets) || !d->w_offsets) return fail(nullptr, DYNOED_EINVAL, "layout offsets missing");
    if (d->n_var <= 0 || d->tau_offset < 0 || d->tau_offset >= d->n_var)
        return fail(nullptr, DYNOED_EINVAL, "bad n_var / tau_offset");
    if (d->substeps <= 0 && !(d->edge_count && d->step_edges))
        return fail(nullptr, DYNOED_EINVAL, "need substeps > 0 or explicit step edges");
    // designs per CTA: 128 unless the caller says otherwise; long design vectors (many intervals /
    // variables) fall back to the largest multiple of 32 whose rows fit the 227 KB of shared memory
    int tile = d->tile ? d->tile : 128;
    if (!d->tile)
        while (tile > 32 && 128 + (size_t)tile * d->n_var * sizeof(double) > (size_t)227 * 1024) tile -= 32;
    // four-stage Rosenbrock gradient with unknown initial conditions: 64-lane CTAs (ic_stride, dynoed_rosenbrock.cuh)
    const bool ic2_four_stage = d->method == DYNOED_RODAS3 && mi.n_ic > 0 && mi.n_ctrl == 0 && m->ic2_ok &&
                                !getenv("DYNOED_NO_IC2");
    if (!d->tile && ic2_four_stage) tile = std::min(tile, 64);
    if (tile % 32 != 0 || tile < 32 || tile > kMaxTile) return fail(nullptr, DYNOED_EINVAL, "tile must be 32..128, multiple of 32");
    const int nvars = mi.n_ctrl + mi.n_obs;
    // per-variable block lengths from the indicators; all indices must be valid
    std::vector<int> len(nvars, 0);
    for (int v = 0; v < nvars; ++v)
        for (int j = 0; j < N; ++j) {
            const int k = d->indicators[v * N + j];
            if (k < 0) return fail(nullptr, DYNOED_EINVAL, "indicator < 0: a merged interval is not covered by a variable grid");
            len[v] = std::max(len[v], k + 1);
        }
    for (int v = 0; v < nvars; ++v) {
        const int off = v < mi.n_ctrl ? d->ctrl_offsets[v] : d->w_offsets[v - mi.n_ctrl];
        if (off < 0 || off + len[v] > d->n_var) return fail(nullptr, DYNOED_EINVAL, "design layout exceeds n_var");
    }
    std::vector<int> w_len_tmp(len.begin() + mi.n_ctrl, len.end());
    if (mi.n_ic && (d->ic_offset < 0 || d->ic_offset + mi.n_ic > d->n_var))
        return fail(nullptr, DYNOED_EINVAL, "ic_offset out of range");
    for (int j = 0; j < N; ++j)
        if (!(d->tgrid[j + 1] > d->tgrid[j])) return fail(nullptr, DYNOED_EINVAL, "tgrid must be strictly increasing");

    ctx = new dynoed_ctx();
    ctx->device = device;
    ctx->model = m;
    ctx->method = d->method;
    ctx->stages = d->method == DYNOED_RK4 ? TabRK4::S : d->method == DYNOED_TSIT5 ? TabTsit5::S
                : d->method == DYNOED_ROS2 ? TabROS2::S : d->method == DYNOED_ROS3P ? TabROS3P::S : TabRODAS3::S;
    ctx->rosenbrock = rosenbrock;
    ctx->criterion = d->criterion;
    ctx->n_intervals = N;
    ctx->n_var = d->n_var;
    ctx->tile = tile;
    for (int i = 0; i < mi.n_obs; ++i) ctx->w_len[i] = w_len_tmp[i];

    // step tables: t and h per step, same arithmetic as the oracle.  Uniform substeps: h = L/m per
    // merged interval — unless the merged grid is uniform up to rounding (all L_j/m within 64 ulp
    // of hbar = (t_N - t_0)/(N m)); then every step uses the single constant dt = hbar, as a
    // fixed-dt integrator run would (grid points such as k/10 are rounded decimals, and their
    // differences scatter over a few ulp).  Explicit edges: t = t0 + e_s*L, h = (t0 + e_{s+1}*L) - t.
    std::vector<int> first(N + 1, 0);
    std::vector<double> st, sh;
    bool single_h = false;
    {
        double hbar = 0.0;
        if (d->substeps > 0) {
            hbar = (d->tgrid[N] - d->tgrid[0]) / ((double)N * d->substeps);
            single_h = true;
            for (int j = 0; j < N; ++j) {
                const double h = (d->tgrid[j + 1] - d->tgrid[j]) / d->substeps;
                if (!(std::fabs(h - hbar) <= 64.0 * 2.220446049250313e-16 * hbar)) { single_h = false; break; }
            }
        }
        size_t eoff = 0;
        for (int j = 0; j < N; ++j) {
            const double t0 = d->tgrid[j], L = d->tgrid[j + 1] - d->tgrid[j];
            first[j] = (int)st.size();
            if (d->substeps > 0) {
                const double h = single_h ? hbar : L / d->substeps;
                for (int s = 0; s < d->substeps; ++s) { st.push_back(t0 + s * h); sh.push_back(h); }
            } else {
                const int ne = d->edge_count[j];
                if (ne < 2) { delete ctx; return fail(nullptr, DYNOED_EINVAL, "edge_count < 2"); }
                for (int s = 0; s + 1 < ne; ++s) {
                    const double ts = t0 + d->step_edges[eoff + s] * L;
                    st.push_back(ts);
                    sh.push_back((t0 + d->step_edges[eoff + s + 1] * L) - ts);
                }
                eoff += ne;
            }
        }
        first[N] = (int)st.size();
    }
    ctx->n_steps = (int)st.size();
    // one step size for the whole grid => the Runge-Kutta factors h*a, h*b are kernel-parameter
    // constants with compile-time addresses (constant-bank operands of the DFMAs)
    ctx->ucoef = single_h && !rosenbrock && !getenv("DYNOED_NO_UCOEF");

    auto bail = [&](int code, const std::string& msg) { dynoed_destroy(ctx); return fail(nullptr, code, msg); };
#define CUC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bail(DYNOED_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    CUC(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CUC(cudaMalloc(&ctx->d_first_step, sizeof(int) * (N + 1)));
    // one padding entry: the small-batch kernel fetches (t, h) of step s + 1 while it computes step s
    CUC(cudaMalloc(&ctx->d_step_t, sizeof(double) * (st.size() + 1)));
    CUC(cudaMalloc(&ctx->d_step_h, sizeof(double) * (st.size() + 1)));
    CUC(cudaMemset(ctx->d_step_t, 0, sizeof(double) * (st.size() + 1)));
    CUC(cudaMemset(ctx->d_step_h, 0, sizeof(double) * (st.size() + 1)));
    CUC(cudaMalloc(&ctx->d_step_ih, sizeof(double) * (st.size() + 1)));
    CUC(cudaMemset(ctx->d_step_ih, 0, sizeof(double) * (st.size() + 1)));
    CUC(cudaMalloc(&ctx->d_indicators, sizeof(int) * nvars * N));
    CUC(cudaMemcpy(ctx->d_first_step, first.data(), sizeof(int) * (N + 1), cudaMemcpyHostToDevice));
    CUC(cudaMemcpy(ctx->d_step_t, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
    CUC(cudaMemcpy(ctx->d_step_h, sh.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
    {
        std::vector<double> sih(sh.size());
        for (size_t k = 0; k < sh.size(); ++k) sih[k] = 1.0 / sh[k];
        CUC(cudaMemcpy(ctx->d_step_ih, sih.data(), sizeof(double) * sih.size(), cudaMemcpyHostToDevice));
    }
    CUC(cudaMemcpy(ctx->d_indicators, d->indicators, sizeof(int) * nvars * N, cudaMemcpyHostToDevice));
    CUC(cudaMalloc(&ctx->d_best, 4 * sizeof(long long)));
    CUC(cudaMalloc(&ctx->d_ring, 2 * kRing * sizeof(long long)));
    CUC(cudaMemset(ctx->d_ring, 0xff, 2 * kRing * sizeof(long long)));
    CUC(cudaMalloc(&ctx->d_ticket, 2 * sizeof(unsigned int)));
    CUC(cudaMemset(ctx->d_ticket, 0, 2 * sizeof(unsigned int)));
    CUC(cudaMalloc(&ctx->d_guard, 8 * sizeof(unsigned long long)));
    CUC(cudaMemset(ctx->d_guard, 0, 8 * sizeof(unsigned long long)));
    ctx->h_first = first;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    if (const char* hf = getenv("DYNOED_L2_HOT_FRAC")) ctx->l2_hot_frac = std::min(4.0, std::max(0.0, atof(hf)));
    ctx->pdl_enabled = !getenv("DYNOED_NO_PDL");
    {   // direction-parallel small-batch kernel: its quadrature scratch must fit what a launch owns anyway
        int nvals = 0;
        for (int k = 0; k < mi.n_ctrl; ++k) nvals += len[k];
        ctx->dir_nl = 1 + mi.n_ic + nvals;
        // measured on B200 (scripts/small_batch_probe.py): ahead of the adjoint kernel up to ~3 warps per SM
        ctx->dir_max_lanes = (int64_t)prop.multiProcessorCount * 3 * kDirBlock;
        if (const char* dl = getenv("DYNOED_DIRPAR_MAX_LANES")) ctx->dir_max_lanes = std::max(0, atoi(dl));
        ctx->dir_ok = !getenv("DYNOED_NO_DIRPAR") && ctx->dir_nl > 1 &&
                      (rosenbrock || (size_t)N * mi.n_obs * mi.nF <= (size_t)ctx->n_steps * mi.nz);
        // time-parallel kernel: one CTA per design, all of its tables in (opt-in) shared memory
        // measured on B200 (scripts/tp_probe.py): ahead of the direction-parallel kernel for every model while
        // each design has an SM of its own; at two CTAs per SM the cheap models lose
        ctx->tp_max_n = (int64_t)prop.multiProcessorCount;
        if (const char* tn = getenv("DYNOED_TIMEPAR_MAX_N")) ctx->tp_max_n = std::max(0, atoi(tn));
        const bool fits = m->tp_smem(d->n_var, N, ctx->n_steps, nvals, false) + 2048 <= (size_t)prop.sharedMemPerBlockOptin;
        ctx->tp_spill = !fits || getenv("DYNOED_TIMEPAR_SPILL");
        ctx->tp_ok = ctx->dir_ok && !getenv("DYNOED_NO_TIMEPAR") && m->tp_nd > 0 &&
                     (fits || m->tp_smem(d->n_var, N, ctx->n_steps, nvals, true) + 2048 <= (size_t)prop.sharedMemPerBlockOptin);
    }
    ctx->zero_copy = !getenv("DYNOED_NO_ZEROCOPY");
    ctx->host_staging = !getenv("DYNOED_NO_STAGING");
    if (const char* hc = getenv("DYNOED_HOST_CHUNKS")) ctx->host_chunks = std::max(0, atoi(hc));

    EvalParams& P = ctx->base;
    memset(&P, 0, sizeof(P));
    P.n_var = d->n_var;
    P.tile = tile;
    P.n_intervals = N;
    P.n_steps = ctx->n_steps;
    P.first_step = ctx->d_first_step;
    P.step_t = ctx->d_step_t;
    P.step_h = ctx->d_step_h;
    P.step_ih = ctx->d_step_ih;
    P.indicators = ctx->d_indicators;
    for (int k = 0; k < mi.n_ctrl; ++k) { P.ctrl_off[k] = d->ctrl_offsets[k]; P.ctrl_len[k] = len[k]; P.n_ctrl_values += len[k]; }
    for (int k = 0; k < mi.n_obs; ++k) P.w_off[k] = d->w_offsets[k];
    P.ic_off = d->ic_offset;
    P.tau_off = d->tau_offset;
    P.criterion = d->criterion;
    for (int k = 0; k < mi.n_theta; ++k) P.theta[k] = d->theta ? d->theta[k] : mi.theta[k];
    for (int k = 0; k < mi.nx; ++k) P.x0[k] = d->x0 ? d->x0[k] : mi.x0[k];
    for (int k = 0; k < mi.nx * mi.np; ++k) P.G0[k] = d->G0 ? d->G0[k] : mi.G0[k];
    if (ctx->ucoef) {
        if (d->method == DYNOED_RK4) fill_coef<TabRK4>(sh[0], P.coef);
        else fill_coef<TabTsit5>(sh[0], P.coef);
    }

    if (const char* cap = getenv("DYNOED_MAX_CTAS_PER_SM")) {   // tuning knob (occupancy experiments)
        const int k = atoi(cap);
        if (k > 0) ctx->smem_floor = ((size_t)prop.sharedMemPerMultiprocessor / k - 1024) & ~(size_t)127;
        if (ctx->smem_floor > (size_t)prop.sharedMemPerBlockOptin - 2048) ctx->smem_floor = prop.sharedMemPerBlockOptin - 2048;
    }
    // Rosenbrock gradient with unknown initial conditions and no controls: second-order sensitivity kernel,
    // unless its per-lane shared-memory columns do not fit (then: forward-mode duals, ros_eval_kernel)
    ctx->ic2 = rosenbrock && mi.n_ic > 0 && mi.n_ctrl == 0 && m->ic2_ok && !getenv("DYNOED_NO_IC2");
    if (ctx->ic2 && ctx->stages >= 4 && ctx->tile > 64) ctx->ic2 = false;     // caller-forced tile above the slot stride
    if (ctx->ic2 && smem_bytes(ctx, 1) > (size_t)prop.sharedMemPerBlockOptin) ctx->ic2 = false;
    // the kernel takes d x / d ic_a from G column ic2_gcol(a): only valid while that column starts as the
    // unit vector of the state (a caller-supplied G0 may say otherwise)
    for (int a = 0; ctx->ic2 && a < mi.n_ic; ++a)
        for (int i = 0; i < mi.nx; ++i)
            if (P.G0[m->ic2_gcol[a] * mi.nx + i] != (i == mi.ic_state[a] ? 1.0 : 0.0)) ctx->ic2 = false;
    if (smem_bytes(ctx) > (size_t)prop.sharedMemPerBlockOptin)
        return bail(DYNOED_EINVAL, "tile * n_var does not fit in shared memory; use a smaller tile");
    for (int g = 0; g < 3; ++g) {
        const size_t smem = smem_bytes(ctx, g);
        CUC(m->attrs(method_code(ctx, g), g >= 1, ctx->ucoef, tile, smem, &ctx->fa[g], &ctx->bps[g]));
        if (ctx->bps[g] < 1) return bail(DYNOED_ECUDA, "kernel cannot be resident with this tile size");
    }
#undef CUC
    *out = ctx;
    return DYNOED_OK;
}

void dynoed_destroy(dynoed_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();   // launches may still be running on a caller-owned stream: scratch is freed below
    cudaGetLastError();
    for (auto& s : ctx->slots) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        cudaFree(s.d_designs); cudaFree(s.d_crit); cudaFree(s.d_grad); cudaFree(s.d_fim); cudaFree(s.traj);
        cudaFreeHost(s.h_designs); cudaFreeHost(s.h_crit); cudaFreeHost(s.h_grad); cudaFreeHost(s.h_fim);
        if (s.done) cudaEventDestroy(s.done);
    }
    cudaFree(ctx->d_first_step); cudaFree(ctx->d_step_t); cudaFree(ctx->d_step_h); cudaFree(ctx->d_step_ih); cudaFree(ctx->d_indicators);
    for (int q = 0; q < 2; ++q) { cudaFree(ctx->traj[q]); cudaFree(ctx->part_val[q]); cudaFree(ctx->part_idx[q]); if (ctx->call_done[q]) cudaEventDestroy(ctx->call_done[q]); }
    cudaFree(ctx->d_best); cudaFree(ctx->d_ring); cudaFree(ctx->d_ticket); cudaFree(ctx->d_guard);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    cudaFree(ctx->hs.pert); cudaFree(ctx->hs.pcrit); cudaFree(ctx->hs.pgrad); cudaFree(ctx->hs.x); cudaFree(ctx->hs.c); cudaFree(ctx->hs.g); cudaFree(ctx->hs.H);
    if (ctx->ev_k0) { cudaEventDestroy(ctx->ev_k0); cudaEventDestroy(ctx->ev_k1); }
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int dynoed_set_criterion(dynoed_ctx* ctx, int criterion) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (criterion < 0 || criterion >= kNumCriteria) return fail(ctx, DYNOED_EINVAL, "criterion out of range");
    ctx->criterion = criterion;
    ctx->base.criterion = criterion;
    return DYNOED_OK;
}

int dynoed_set_stream(dynoed_ctx* ctx, void* s) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (ctx->stream == (cudaStream_t)s) return DYNOED_OK;   // unchanged: keeps the overlap bookkeeping
    // the ctx's scratch is ordered by its stream: drain the old one before work moves to another
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)s;   // NULL is the CUDA legacy default stream
    ctx->prev.valid = false;
    return DYNOED_OK;
}

int dynoed_own_stream(dynoed_ctx* ctx) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (ctx->stream != ctx->own_stream) { CU(cudaSetDevice(ctx->device)); CU(cudaStreamSynchronize(ctx->stream)); }
    ctx->stream = ctx->own_stream;
    ctx->prev.valid = false;
    return DYNOED_OK;
}

// cudaPointerGetAttributes, memoised for the duration of ONE evaluation call (the same four caller
// pointers are classified several times on the way down; a one-design call is all host overhead)
struct AttrMemo {
    bool active = false;
    int n = 0;
    const void* ptr[8];
    cudaPointerAttributes attr[8];
    bool ok[8];
};
static thread_local AttrMemo g_attr_memo;
struct AttrMemoScope {
    AttrMemoScope() { g_attr_memo.active = true; g_attr_memo.n = 0; }
    ~AttrMemoScope() { g_attr_memo.active = false; g_attr_memo.n = 0; }
};
static bool pointer_attrs(const void* p, cudaPointerAttributes* out) {
    AttrMemo& m = g_attr_memo;
    if (m.active)
        for (int i = 0; i < m.n; ++i)
            if (m.ptr[i] == p) { *out = m.attr[i]; return m.ok[i]; }
    const bool ok = cudaPointerGetAttributes(out, p) == cudaSuccess;
    if (!ok) cudaGetLastError();
    if (m.active && m.n < 8) { m.ptr[m.n] = p; m.attr[m.n] = *out; m.ok[m.n] = ok; ++m.n; }
    return ok;
}

// 0 = host, 1 = device/managed
static int ptr_kind(const void* p) {
    cudaPointerAttributes a;
    if (!pointer_attrs(p, &a)) return 0;
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
}

static bool ranges_overlap(const char* a0, const char* a1, const char* b0, const char* b1) {
    return a0 && b0 && a0 < b1 && b0 < a1;
}

// How much of the checkpoint stack stays in L2 (see the CkptPolicy comment in dynoed_kernels.cuh):
// the top `hot` steps of every resident lane, sized so that all resident lanes together use
// l2_hot_frac of the L2.
static void set_scratch_params(dynoed_ctx* ctx, EvalParams& P, int par, int kidx) {
    const auto& mi = ctx->model->info;
    P.guard = ctx->d_guard + 4 * par;
    P.serial = ctx->set_calls[par] - 1;      // the calls that used this set before the current one must be done
    const double lanes = (double)ctx->sm_count * std::max(ctx->bps[kidx], 1) * kMaxTile;
    const double per_step = 8.0 * (mi.nz + (double)mi.n_obs * mi.nF * ctx->n_intervals / std::max(ctx->n_steps, 1));
    int hot = (int)((double)ctx->l2_bytes * ctx->l2_hot_frac / (lanes * per_step));
    hot = std::max(0, std::min(hot, ctx->n_steps));
    P.hot_step = ctx->n_steps - hot;
    P.hot_interval = ctx->n_intervals;
    for (int j = 0; j < ctx->n_intervals; ++j)
        if (ctx->h_first[j] >= P.hot_step) { P.hot_interval = j; break; }
}

// after a failed call the scratch set may be left acquired / the ticket partially counted: drain and clear
static void reset_parity(dynoed_ctx* ctx, int par) {
    cudaDeviceSynchronize();
    cudaGetLastError();
    cudaMemset(ctx->d_ticket + par, 0, sizeof(unsigned int));
    cudaMemcpy(ctx->d_guard + 4 * par, &ctx->set_calls[par], sizeof(unsigned long long), cudaMemcpyHostToDevice);   // all done
    ctx->prev.valid = false;
}

// launch on device-resident buffers
static int launch_device(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad,
                         double* fim, double* xgrid, int par, int grid, cudaStream_t st) {
    const auto& mi = ctx->model->info;
    EvalParams P = ctx->base;
    P.designs = designs; P.crit = crit; P.grad = grad; P.fim = fim; P.xgrid = xgrid;
    P.traj = ctx->traj[par]; P.n = n;
    P.part_val = ctx->part_val[par]; P.part_idx = ctx->part_idx[par];
    P.part_base = 0; P.n_parts = grid; P.ticket_target = (unsigned int)grid;
    P.ticket = ctx->d_ticket + par; P.best_val = reinterpret_cast<double*>(ctx->d_best + 2 * par); P.best_idx = ctx->d_best + 2 * par + 1;
    P.ring_rec = ctx->d_ring + 2 * (ctx->batch_seq % kRing);
    set_scratch_params(ctx, P, par, grad ? ctx->kidx_req : 0);
    P.use_tma = ((reinterpret_cast<uintptr_t>(designs) % 16 == 0) &&
                 (grad == nullptr || reinterpret_cast<uintptr_t>(grad) % 16 == 0) &&
                 (((size_t)ctx->tile * ctx->n_var * 8) % 16 == 0)) ? 1 : 0;
    // May this launch start while the previous one drains (programmatic dependent launch)?  Only if
    // the caller opted in (DYNOED_OVERLAP_OK: no kernel enqueued after the previous launch produces
    // this launch's inputs — the kernels do not execute griddepcontrol.wait), both launches are the
    // same kernel of this context on the same stream, and their buffers are disjoint.  The two
    // scratch sets alternate; the scratch-set guard (scratch_acquire) keeps launch k+2 off the set
    // of launch k until k has finished, whatever the grid sizes.
    dynoed_ctx::Prev cur;
    cur.kidx = grad ? ctx->kidx_req : 0;
    const bool dirpar = grad && !xgrid && dir_eligible(ctx, n, cur.kidx);
    const bool tpobj = !grad && !xgrid && tp_obj_eligible(ctx, n, cur.kidx);
    const bool timepar = (dirpar && tp_eligible(ctx, n, cur.kidx)) || tpobj;
    cur.valid = xgrid == nullptr && !dirpar && !tpobj;
    if (dirpar || tpobj) P.use_tma = 0;
    cur.full = !dirpar && !tpobj && grid == ctx->sm_count * std::max(ctx->bps[cur.kidx], 1);
    cur.stream = st;
    cur.in0 = (const char*)designs; cur.in1 = cur.in0 + sizeof(double) * n * ctx->n_var;
    cur.out0[0] = (const char*)crit; cur.out1[0] = cur.out0[0] + sizeof(double) * n;
    cur.out0[1] = (const char*)grad; cur.out1[1] = grad ? cur.out0[1] + sizeof(double) * n * ctx->n_var : nullptr;
    cur.out0[2] = (const char*)fim; cur.out1[2] = fim ? cur.out0[2] + sizeof(double) * n * mi.nF : nullptr;
    const dynoed_ctx::Prev& pv = ctx->prev;
    bool pdl = ctx->pdl_enabled && (ctx->call_flags & DYNOED_OVERLAP_OK) && !ctx->profiling && cur.valid && pv.valid &&
               cur.kidx == pv.kidx && cur.stream == pv.stream && stream_last_serial(st) == pv.serial;
    for (int a = 0; pdl && a < 3; ++a) {
        if (ranges_overlap(cur.in0, cur.in1, pv.out0[a], pv.out1[a])) pdl = false;      // RAW
        if (ranges_overlap(cur.out0[a], cur.out1[a], pv.in0, pv.in1)) pdl = false;      // WAR
        for (int b = 0; pdl && b < 3; ++b)
            if (ranges_overlap(cur.out0[a], cur.out1[a], pv.out0[b], pv.out1[b])) pdl = false;   // WAW
    }
    if (ctx->profiling) {
        if (ctx->ev_pending) {   // fold the previous launch's duration into the running sum
            CU(cudaEventSynchronize(ctx->ev_k1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1));
            ctx->kernel_ms_sum += ms; ctx->kernel_ms_count++;
            ctx->ev_pending = false;
        }
        CU(cudaEventRecord(ctx->ev_k0, st));
    }
    CU(ctx->model->launch(method_code(ctx, cur.kidx) | (timepar ? (ctx->tp_spill ? kTimePar | kTpSpill : kTimePar) : dirpar ? kDirPar : 0), grad != nullptr,
                          ctx->ucoef, pdl, P, grid, ctx->tile, smem_bytes(ctx, cur.kidx), st));
    if (ctx->profiling) { CU(cudaEventRecord(ctx->ev_k1, st)); ctx->ev_pending = true; }
    cur.serial = stream_new_serial(st);
    ctx->prev = cur;
    ctx->stream_dirty = true;
    ctx->launches++;
    ctx->pdl_launches += pdl ? 1 : 0;
    ctx->dir_launches += (dirpar || tpobj) ? 1 : 0;
    ctx->tp_launches += timepar ? 1 : 0;
    return DYNOED_OK;
}

static int eval_device(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad, double* fim,
                       double* xgrid) {
    const bool g = grad != nullptr;
    if (g && xgrid) return fail(ctx, DYNOED_EUNSUPPORTED, "the trajectory is written by the objective-only kernels");
    const int kidx = g ? ctx->kidx_req : 0;
    ctx->call_n = n;
    ctx->call_xgrid = xgrid != nullptr;
    const int grid = grid_for(ctx, n, kidx);
    const int par = ctx->parity;
    if (g) {
        const size_t need = traj_need(ctx, n, kidx);
        if (need > ctx->traj_bytes[par]) {
            if (ctx->traj[par]) { CU(cudaDeviceSynchronize()); cudaFree(ctx->traj[par]); ctx->traj[par] = nullptr; }
            const size_t full = traj_bytes_for(ctx, ctx->sm_count * std::max(ctx->bps[kidx], 1), kidx);
            const size_t want = std::max(need, std::min(full, need * 4));
            CU(cudaMalloc(&ctx->traj[par], want));
            ctx->scratch_bytes += want - ctx->traj_bytes[par];
            ctx->traj_bytes[par] = want;
        }
    }
    int rc = ensure_partials(ctx, grid, par);
    if (rc) return rc;
    ctx->call_serial++;
    ctx->set_calls[par]++;
    rc = launch_device(ctx, designs, n, crit, grad, fim, xgrid, par, grid, ctx->stream);
    if (rc) { reset_parity(ctx, par); return rc; }
    ctx->last_parity = par;
    ctx->parity = par ^ 1;
    ctx->batch_seq++;
    ctx->have_batch = true;
    return DYNOED_OK;
}

static int ensure_slot(dynoed_ctx* ctx, Slot& s, int64_t cap, int kidx, bool fim) {
    const bool grad = kidx > 0;
    const int nF = ctx->model->info.nF;
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (cap > s.cap) {
        CU(cudaStreamSynchronize(s.stream));
        cudaFree(s.d_designs); cudaFree(s.d_crit); cudaFree(s.d_grad); cudaFree(s.d_fim);
        s.d_designs = s.d_crit = s.d_grad = s.d_fim = nullptr;
        CU(cudaMalloc(&s.d_designs, sizeof(double) * cap * ctx->n_var));
        CU(cudaMalloc(&s.d_crit, sizeof(double) * cap));
        CU(cudaMalloc(&s.d_grad, sizeof(double) * cap * ctx->n_var));
        CU(cudaMalloc(&s.d_fim, sizeof(double) * cap * nF));
        ctx->scratch_bytes += sizeof(double) * (size_t)(cap - s.cap) * (2 * ctx->n_var + 1 + nF);
        s.cap = cap;
    }
    if (grad) {
        const size_t need = traj_need(ctx, cap, kidx);
        if (need > s.traj_bytes) {
            CU(cudaStreamSynchronize(s.stream));
            cudaFree(s.traj); s.traj = nullptr;
            CU(cudaMalloc(&s.traj, need));
            ctx->scratch_bytes += need - s.traj_bytes;
            s.traj_bytes = need;
        }
    }
    (void)fim;
    return DYNOED_OK;
}

static int ensure_staging(dynoed_ctx* ctx, Slot& s, int64_t cap) {
    const int nF = ctx->model->info.nF;
    if (!s.done) CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (cap > s.hcap) {
        CU(cudaStreamSynchronize(s.stream));
        cudaFreeHost(s.h_designs); cudaFreeHost(s.h_crit); cudaFreeHost(s.h_grad); cudaFreeHost(s.h_fim);
        s.h_designs = s.h_crit = s.h_grad = s.h_fim = nullptr;
        CU(cudaMallocHost(&s.h_designs, sizeof(double) * cap * ctx->n_var));
        CU(cudaMallocHost(&s.h_crit, sizeof(double) * cap));
        CU(cudaMallocHost(&s.h_grad, sizeof(double) * cap * ctx->n_var));
        CU(cudaMallocHost(&s.h_fim, sizeof(double) * cap * nF));
        // device addresses of the staging, looked up once (the small zero-copy calls would pay four driver
        // calls per evaluation otherwise)
        CU(cudaHostGetDevicePointer((void**)&s.hd_designs, s.h_designs, 0));
        CU(cudaHostGetDevicePointer((void**)&s.hd_crit, s.h_crit, 0));
        CU(cudaHostGetDevicePointer((void**)&s.hd_grad, s.h_grad, 0));
        CU(cudaHostGetDevicePointer((void**)&s.hd_fim, s.h_fim, 0));
        s.hcap = cap;
    }
    return DYNOED_OK;
}

// is this host pointer ordinary pageable memory (not page-locked / registered)?
static bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (!pointer_attrs(p, &a)) return true;
    return a.type == cudaMemoryTypeUnregistered;
}

// host buffers: chunked pipeline, chunk k on stream k % kSlots (H2D -> kernel -> D2H in order on
// that stream; different chunks overlap copy engines and SMs)
static int eval_host(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad, double* fim,
                     bool pipelined = false) {
    const int nF = ctx->model->info.nF;
    const bool g = grad != nullptr;
    ctx->call_n = n;
    ctx->call_xgrid = false;
    // Chunk schedule: host_chunks equal chunks (default 8).  Measured on B200 / PCIe Gen5: 4..16 equal
    // chunks and a ramped schedule (small chunks at both ends: 1,1,2,4,8,8,4,2,1,1 thirty-seconds,
    // DYNOED_HOST_CHUNKS=0) all land within 5 % of each other at 71-75 % of the rate of two
    // concurrent pinned copies: the call is bound by the copy engines, not by fill / drain.
    const int64_t tile = ctx->tile;
    const int64_t ntiles = (n + tile - 1) / tile;
    std::vector<int64_t> bounds{0};
    if (ctx->host_chunks > 0 || ntiles < 64) {
        const int64_t k = ctx->host_chunks > 0 ? ctx->host_chunks : 8;
        int64_t chunk = std::max<int64_t>(((ntiles + k - 1) / k), std::min<int64_t>(16, ntiles)) * tile;
        chunk = std::min<int64_t>(chunk, 65536);   // bounds the per-slot device and staging buffers for huge batches
        for (int64_t o = chunk; o < n; o += chunk) bounds.push_back(o);
    } else {
        static const int ramp[10] = {1, 1, 2, 4, 8, 8, 4, 2, 1, 1};
        int64_t done = 0;
        for (int r = 0; r < 9; ++r) {
            done += ramp[r];
            bounds.push_back(std::min<int64_t>((ntiles * done / 32) * tile, n));
        }
    }
    bounds.push_back(n);
    bounds.erase(std::unique(bounds.begin(), bounds.end()), bounds.end());
    const int64_t nchunks = (int64_t)bounds.size() - 1;
    int64_t chunk = 0;
    for (int64_t c = 0; c < nchunks; ++c) chunk = std::max(chunk, bounds[c + 1] - bounds[c]);
    int total_parts = 0;
    const int kidx = g ? ctx->kidx_req : 0;
    for (int64_t c = 0; c < nchunks; ++c) total_parts += grid_for(ctx, bounds[c + 1] - bounds[c], kidx);
    const int par = ctx->parity;
    ctx->prev.valid = false;    // host-path launches run on their own streams: no overlap bookkeeping
    int rc = ensure_partials(ctx, total_parts, par);
    if (rc) return rc;
    for (int k = 0; k < kSlots && k < nchunks; ++k) {
        rc = ensure_slot(ctx, ctx->slots[k], chunk, kidx, fim != nullptr);
        if (rc) return rc;
    }
    // Pinned (page-locked, mapped) result arrays: the kernel stores crit / fim straight into host
    // memory (posted PCIe writes, 8 + 8 nF bytes per design), which removes two of the three
    // device-to-host copies per chunk; the gradient still travels as one bulk copy.
    auto mapped = [ctx](const void* p) -> double* {
        if (!p || !ctx->zero_copy) return nullptr;
        cudaPointerAttributes a;
        if (!pointer_attrs(p, &a)) return nullptr;
        return a.type == cudaMemoryTypeHost ? static_cast<double*>(a.devicePointer) : nullptr;
    };
    double* crit_map = mapped(crit);
    double* fim_map = mapped(fim);
    // gradient tiles stored by the kernel's TMA bulk store straight into page-locked host memory (no D2H copy
    // stage in the pipeline): opt-in, DYNOED_ZC_GRAD=1 (measurement switch)
    static const bool zc_grad_on = getenv("DYNOED_ZC_GRAD") && atoi(getenv("DYNOED_ZC_GRAD")) != 0;
    double* grad_map = (g && zc_grad_on) ? mapped(grad) : nullptr;
    // Pageable caller arrays: the driver would stage them through a small bounce buffer at a few
    // GB/s.  Instead every chunk goes through page-locked staging owned by its slot, filled and
    // drained by multi-threaded memcpy while the other slots' copies and kernels are in flight.
    const bool staged = ctx->host_staging &&
                        (is_pageable(designs) || is_pageable(crit) || is_pageable(grad) || is_pageable(fim));
    if (staged) {
        crit_map = fim_map = grad_map = nullptr;
        for (int k = 0; k < kSlots && k < nchunks; ++k) {
            rc = ensure_staging(ctx, ctx->slots[k], chunk);
            if (rc) return rc;
        }
    }
    auto drain = [&](int64_t c) -> int {     // chunk c has landed in its slot's staging: hand it to the caller
        Slot& s = ctx->slots[c % kSlots];
        const int64_t off = bounds[c], cnt = bounds[c + 1] - bounds[c];
        CU(cudaEventSynchronize(s.done));
        memcpy(crit + off, s.h_crit, sizeof(double) * cnt);
        if (g) parallel_memcpy(grad + off * ctx->n_var, s.h_grad, sizeof(double) * cnt * ctx->n_var);
        if (fim) memcpy(fim + off * nF, s.h_fim, sizeof(double) * cnt * nF);
        return DYNOED_OK;
    };
    if (!ctx->ev_order) CU(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
    cudaEvent_t ev = ctx->ev_order;   // re-recorded at every use: a wait captures the record that precedes it
    for (int q = 0; q < 2; ++q)
        if (!ctx->call_done[q]) CU(cudaEventCreateWithFlags(&ctx->call_done[q], cudaEventDisableTiming));
    // A call of a handful of designs (the optimiser callback) is all latency: skip the three or four
    // small DMA transfers — the small-batch kernels stage the rows from page-locked host
    // memory themselves (one coalesced read) and store their results there (posted writes).  Such a call
    // runs on the ctx stream itself: no slot stream, no cross-stream events (four driver calls less).
    bool small_zc = false;
    double *zc_in = nullptr, *zc_crit = nullptr, *zc_grad = nullptr, *zc_fim = nullptr;
    if (!g && nchunks == 1 && n <= kSmallZeroCopy && ctx->zero_copy && !pipelined && tp_obj_eligible(ctx, n, kidx)) {
        Slot& s0 = ctx->slots[0];
        auto dev = [](const void* p) -> double* {
            cudaPointerAttributes a;
            if (!p || !pointer_attrs(p, &a)) return nullptr;
            return a.type == cudaMemoryTypeHost ? static_cast<double*>(a.devicePointer) : nullptr;
        };
        zc_in = staged ? s0.hd_designs : dev(designs);
        zc_crit = staged ? s0.hd_crit : dev(crit);
        zc_fim = !fim ? nullptr : staged ? s0.hd_fim : dev(fim);
        small_zc = zc_in && zc_crit && (!fim || zc_fim);
    }
    if (g && nchunks == 1 && n <= kSmallZeroCopy && ctx->zero_copy && !pipelined && dir_eligible(ctx, n, kidx) &&
        (tp_eligible(ctx, n, kidx) || ((size_t)(kDirBlock + ctx->dir_nl - 1) / ctx->dir_nl + 1) *
                (ctx->n_var + (size_t)ctx->n_intervals * ctx->model->info.n_obs * ctx->model->info.nF) * sizeof(double) +
                sizeof(int) * ((size_t)(ctx->model->info.n_ctrl + ctx->model->info.n_obs + 1) * ctx->n_intervals + 1) <= 48 * 1024)) {
        Slot& s0 = ctx->slots[0];
        auto dev = [](const void* p) -> double* {
            cudaPointerAttributes a;
            if (!p || !pointer_attrs(p, &a)) return nullptr;
            return a.type == cudaMemoryTypeHost ? static_cast<double*>(a.devicePointer) : nullptr;
        };
        zc_in = staged ? s0.hd_designs : dev(designs);
        zc_crit = staged ? s0.hd_crit : dev(crit);
        zc_grad = staged ? s0.hd_grad : dev(grad);
        zc_fim = !fim ? nullptr : staged ? s0.hd_fim : dev(fim);
        small_zc = zc_in && zc_crit && zc_grad && (!fim || zc_fim);
    }
    if (pipelined && !ctx->stream_dirty && !staged) {
        // Pipelined host call: do NOT order after the previous host call (its D2H should overlap this
        // call's H2D; chunk c of both calls shares slot c % kSlots, whose stream keeps them in order).
        // Only the call that used the same scratch parity (two calls back) must have finished.
        if (ctx->call_of_parity[par] >= 0)
            for (int k = 0; k < kSlots && k < nchunks; ++k)
                CU(cudaStreamWaitEvent(ctx->slots[k].stream, ctx->call_done[par], 0));
    } else if (!small_zc) {
        // order after whatever the ctx stream was doing
        CU(cudaEventRecord(ev, ctx->stream));
        for (int k = 0; k < kSlots && k < nchunks; ++k) CU(cudaStreamWaitEvent(ctx->slots[k].stream, ev, 0));
    }
    ctx->call_serial++;
    ctx->set_calls[par]++;
    // everything that launches: a failure part-way leaves the ticket partially counted and the scratch set
    // acquired, so the error path drains the device and clears both (reset_parity)
    auto enqueue = [&]() -> int {
        int part_base = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            Slot& s = ctx->slots[c % kSlots];
            const int64_t off = bounds[c], cnt = bounds[c + 1] - bounds[c];
            const double* h_src = designs + off * ctx->n_var;
            if (staged) {
                if (c >= kSlots) { rc = drain(c - kSlots); if (rc) return rc; }   // frees this slot's staging
                if (small_zc) memcpy(s.h_designs, h_src, sizeof(double) * cnt * ctx->n_var);
                else parallel_memcpy(s.h_designs, h_src, sizeof(double) * cnt * ctx->n_var);
                h_src = s.h_designs;
            }
            if (!small_zc)
                CU(cudaMemcpyAsync(s.d_designs, h_src, sizeof(double) * cnt * ctx->n_var, cudaMemcpyHostToDevice, s.stream));
            EvalParams P = ctx->base;
            const int grid = grid_for(ctx, cnt, kidx);
            P.designs = s.d_designs; P.crit = crit_map ? crit_map + off : s.d_crit; P.grad = !g ? nullptr : grad_map ? grad_map + off * ctx->n_var : s.d_grad;
            P.fim = !fim ? nullptr : fim_map ? fim_map + off * nF : s.d_fim;
            if (small_zc) {   // the kernel reads the rows from, and stores every result into, page-locked host memory
                P.designs = zc_in; P.crit = zc_crit; P.grad = g ? zc_grad : nullptr; P.fim = fim ? zc_fim : nullptr;
            }
            P.xgrid = nullptr; P.traj = s.traj; P.n = cnt;
            P.part_val = ctx->part_val[par]; P.part_idx = ctx->part_idx[par];
            P.part_base = part_base; P.n_parts = total_parts; P.ticket_target = (unsigned int)total_parts;
            P.ticket = ctx->d_ticket + par; P.best_val = reinterpret_cast<double*>(ctx->d_best + 2 * par); P.best_idx = ctx->d_best + 2 * par + 1;
            P.ring_rec = ctx->d_ring + 2 * (ctx->batch_seq % kRing);
            set_scratch_params(ctx, P, par, kidx);
            part_base += grid;
            const bool dirpar = g && dir_eligible(ctx, cnt, kidx);
            const bool tpobj = !g && tp_obj_eligible(ctx, cnt, kidx);
            const bool timepar = (dirpar && tp_eligible(ctx, cnt, kidx)) || tpobj;
            P.use_tma = (!dirpar && !tpobj && ((size_t)ctx->tile * ctx->n_var * 8) % 16 == 0) ? 1 : 0;
            P.index_base = off;
            cudaStream_t run_stream = small_zc ? ctx->stream : s.stream;
            CU(ctx->model->launch(method_code(ctx, kidx) | (timepar ? (ctx->tp_spill ? kTimePar | kTpSpill : kTimePar) : dirpar ? kDirPar : 0), g, ctx->ucoef, false, P,
                                  grid, ctx->tile, smem_bytes(ctx, kidx), run_stream));
            ctx->launches++;
            ctx->dir_launches += (dirpar || tpobj) ? 1 : 0;
            ctx->tp_launches += timepar ? 1 : 0;
            double* o_crit = staged ? s.h_crit : crit + off;
            double* o_grad = staged ? s.h_grad : (g ? grad + off * ctx->n_var : nullptr);
            double* o_fim = staged ? s.h_fim : (fim ? fim + off * nF : nullptr);
            if (!crit_map && !small_zc) CU(cudaMemcpyAsync(o_crit, s.d_crit, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s.stream));
            if (g && !small_zc && !grad_map) CU(cudaMemcpyAsync(o_grad, s.d_grad, sizeof(double) * cnt * ctx->n_var, cudaMemcpyDeviceToHost, s.stream));
            if (fim && !fim_map && !small_zc) CU(cudaMemcpyAsync(o_fim, s.d_fim, sizeof(double) * cnt * nF, cudaMemcpyDeviceToHost, s.stream));
            if (staged) CU(cudaEventRecord(s.done, run_stream));
        }
        if (staged)
            for (int64_t c = std::max<int64_t>(0, nchunks - kSlots); c < nchunks; ++c) { rc = drain(c); if (rc) return rc; }
        for (int k = 0; k < kSlots && k < nchunks && !small_zc; ++k) {
            CU(cudaEventRecord(ev, ctx->slots[k].stream));
            CU(cudaStreamWaitEvent(ctx->stream, ev, 0));
        }
        return DYNOED_OK;
    };
    rc = enqueue();
    if (rc) { reset_parity(ctx, par); return rc; }
    CU(cudaEventRecord(ctx->call_done[par], ctx->stream));   // fires when every chunk of THIS call has landed
    ctx->host_seq++;
    ctx->call_of_parity[par] = ctx->host_seq;
    ctx->stream_dirty = false;
    ctx->last_parity = par;
    ctx->parity = par ^ 1;
    ctx->batch_seq++;
    ctx->have_batch = true;
    return DYNOED_OK;
}

static int eval_common(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad, double* fim,
                       bool allow_host, bool sync, uint32_t flags = 0, bool pipelined = false) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    AttrMemoScope memo;
    ctx->err.clear();
    if (n < 0) return fail(ctx, DYNOED_EINVAL, "n < 0");
    if (n == 0) { ctx->have_batch = false; return DYNOED_OK; }
    if (!designs || !crit) return fail(ctx, DYNOED_EINVAL, "designs/crit is NULL");
    ctx->kidx_req = kernel_index(ctx, grad != nullptr, flags);
    ctx->call_flags = flags;
    CU(cudaSetDevice(ctx->device));
    const int kd = ptr_kind(designs);
    if (ptr_kind(crit) != kd || (grad && ptr_kind(grad) != kd) || (fim && ptr_kind(fim) != kd))
        return fail(ctx, DYNOED_EINVAL, "designs/crit/grad/fim must all be host or all be device pointers");
    int rc;
    if (kd == 1) rc = eval_device(ctx, designs, n, crit, grad, fim, nullptr);
    else {
        if (!allow_host) return fail(ctx, DYNOED_EINVAL, "dynoed_eval_batch_async needs device pointers");
        rc = eval_host(ctx, designs, n, crit, grad, fim, pipelined);
    }
    if (rc) return rc;
    if (sync) CU(cudaStreamSynchronize(ctx->stream));
    return DYNOED_OK;
}

int dynoed_eval_batch(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad, double* fim,
                      uint32_t flags) {
    return eval_common(ctx, designs, n, crit, grad, fim, true, true, flags);
}

int dynoed_eval_batch_async(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad,
                            double* fim, uint32_t flags) {
    return eval_common(ctx, designs, n, crit, grad, fim, false, false, flags);
}

int dynoed_eval_batch_host_async(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad,
                                 double* fim, uint32_t flags, int64_t* call_id) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (n <= 0 || !designs || !crit) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    if (ptr_kind(designs) != 0) return fail(ctx, DYNOED_EINVAL, "dynoed_eval_batch_host_async takes host pointers");
    const int rc = eval_common(ctx, designs, n, crit, grad, fim, true, false, flags, true);
    if (rc) return rc;
    if (call_id) *call_id = ctx->host_seq;
    return DYNOED_OK;
}

int dynoed_wait(dynoed_ctx* ctx, int64_t call_id) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (call_id < 1 || call_id > ctx->host_seq) return fail(ctx, DYNOED_EINVAL, "unknown call id");
    CU(cudaSetDevice(ctx->device));
    // the event of the parity this call used (re-recorded by the call two later, which is ordered after it)
    for (int q = 0; q < 2; ++q)
        if (ctx->call_of_parity[q] >= call_id && ((ctx->call_of_parity[q] - call_id) % 2 == 0)) {
            CU(cudaEventSynchronize(ctx->call_done[q]));
            return DYNOED_OK;
        }
    CU(cudaStreamSynchronize(ctx->stream));
    return DYNOED_OK;
}

int dynoed_sync(dynoed_ctx* ctx) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return DYNOED_OK;
}

int dynoed_argmin(dynoed_ctx* ctx, int64_t* idx, double* val) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!ctx->have_batch) return fail(ctx, DYNOED_EINVAL, "no batch has been evaluated");
    CU(cudaSetDevice(ctx->device));
    double v; long long i;
    long long rec[2];
    CU(cudaMemcpyAsync(rec, ctx->d_best + 2 * ctx->last_parity, sizeof(rec), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(&v, &rec[0], sizeof(double));
    i = rec[1];
    if (idx) *idx = (int64_t)i;
    if (val) *val = v;
    return DYNOED_OK;
}

int dynoed_argmin_async(dynoed_ctx* ctx, double* d_val, int64_t* d_idx) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!ctx->have_batch) return fail(ctx, DYNOED_EINVAL, "no batch has been evaluated");
    if (!d_val || !d_idx) return fail(ctx, DYNOED_EINVAL, "destination is NULL");
    CU(cudaMemcpyAsync(d_val, ctx->d_best + 2 * ctx->last_parity, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_idx, ctx->d_best + 2 * ctx->last_parity + 1, sizeof(long long), cudaMemcpyDeviceToDevice, ctx->stream));
    return DYNOED_OK;
}

int dynoed_argmin_record_async(dynoed_ctx* ctx, void* d_record) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!ctx->have_batch) return fail(ctx, DYNOED_EINVAL, "no batch has been evaluated");
    if (!d_record) return fail(ctx, DYNOED_EINVAL, "destination is NULL");
    CU(cudaMemcpyAsync(d_record, ctx->d_best + 2 * ctx->last_parity, 2 * sizeof(long long),
                       cudaMemcpyDeviceToDevice, ctx->stream));
    return DYNOED_OK;
}

int dynoed_set_aux_stream(dynoed_ctx* ctx, void* s) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    ctx->aux_stream = (cudaStream_t)s;
    return DYNOED_OK;
}

// the stream the ring-reading selection helpers run on: the aux stream (ordered after every evaluation
// enqueued so far by an event — an event record does not break the launches' overlap, a kernel or copy
// on the evaluation stream would) or, without one, the ctx stream itself
static int ring_reader_stream(dynoed_ctx* ctx, cudaStream_t* out) {
    *out = ctx->stream;
    if (!ctx->aux_stream || ctx->aux_stream == ctx->stream) return DYNOED_OK;
    if (!ctx->ev_aux) CU(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_aux, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux, 0));
    *out = ctx->aux_stream;
    return DYNOED_OK;
}

int dynoed_argmin_history_async(dynoed_ctx* ctx, int count, void* d_records) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (count < 1 || count > kRing || (int64_t)count > ctx->batch_seq)
        return fail(ctx, DYNOED_EINVAL, "count must be 1..64 and not exceed the batches evaluated so far");
    if (!d_records) return fail(ctx, DYNOED_EINVAL, "destination is NULL");
    const int64_t first = ctx->batch_seq - count;          // oldest batch wanted
    const int s0 = (int)(first % kRing);
    const int n0 = std::min(count, kRing - s0);            // up to the end of the ring, then wrap
    char* dst = reinterpret_cast<char*>(d_records);
    cudaStream_t st;
    int rc = ring_reader_stream(ctx, &st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst, ctx->d_ring + 2 * s0, (size_t)n0 * 16, cudaMemcpyDeviceToDevice, st));
    if (n0 < count)
        CU(cudaMemcpyAsync(dst + (size_t)n0 * 16, ctx->d_ring, (size_t)(count - n0) * 16, cudaMemcpyDeviceToDevice, st));
    if (st == ctx->stream) ctx->prev.valid = false;
    return DYNOED_OK;
}

int dynoed_pack_winner(dynoed_ctx* ctx, int nbatches, const double* designs, const double* grad, int64_t n_local,
                       int64_t batch_stride, int64_t index_base, double* d_payload) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (nbatches < 0 || nbatches > kRing || (int64_t)nbatches > ctx->batch_seq)
        return fail(ctx, DYNOED_EINVAL, "nbatches must be 0..64 and not exceed the batches evaluated so far");
    if (!d_payload || (nbatches > 0 && !designs) || n_local < 0 || batch_stride < 1)
        return fail(ctx, DYNOED_EINVAL, "bad arguments");
    if (ptr_kind(d_payload) != 1 || (designs && ptr_kind(designs) != 1) || (grad && ptr_kind(grad) != 1))
        return fail(ctx, DYNOED_EINVAL, "dynoed_pack_winner needs device pointers");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st;
    int rc = ring_reader_stream(ctx, &st);
    if (rc) return rc;
    const int s0 = (int)((ctx->batch_seq - nbatches) % kRing);
    pack_winner_kernel<<<1, kSelThreads, 0, st>>>(ctx->d_ring, kRing, s0, nbatches, (long long)batch_stride,
                                                  (long long)index_base, (long long)n_local, designs, grad, ctx->n_var, d_payload);
    CU(cudaGetLastError());
    ctx->launches++;
    if (st == ctx->stream) ctx->prev.valid = false;
    return DYNOED_OK;
}

int dynoed_select_winner(dynoed_ctx* ctx, const double* d_gathered, int world, double* d_out) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!d_gathered || !d_out || world < 1 || world > 65536) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    if (ptr_kind(d_gathered) != 1 || ptr_kind(d_out) != 1) return fail(ctx, DYNOED_EINVAL, "dynoed_select_winner needs device pointers");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->aux_stream ? ctx->aux_stream : ctx->stream;
    select_winner_kernel<<<1, kSelThreads, 0, st>>>(d_gathered, world, 2 + 2 * ctx->n_var, d_out);
    CU(cudaGetLastError());
    ctx->launches++;
    if (st == ctx->stream) ctx->prev.valid = false;
    return DYNOED_OK;
}

int dynoed_guard_stats(dynoed_ctx* ctx, int64_t* waits, int64_t* timeouts) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    unsigned long long g[8];
    CU(cudaMemcpy(g, ctx->d_guard, sizeof(g), cudaMemcpyDeviceToHost));   // synchronises with all prior work
    if (waits) *waits = (int64_t)(g[1] + g[5]);
    if (timeouts) *timeouts = (int64_t)(g[2] + g[6]);
    return DYNOED_OK;
}

int dynoed_argmin_gathered(dynoed_ctx* ctx, const void* d_records, int world, int groups, int64_t shard_size,
                           void* d_out) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!d_records || !d_out || world < 1 || world > 1024 || groups < 1 || groups > 65535)
        return fail(ctx, DYNOED_EINVAL, "bad arguments");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->aux_stream ? ctx->aux_stream : ctx->stream;
    argmin_records_kernel<<<groups, 32, 0, st>>>(reinterpret_cast<const long long*>(d_records), world, groups,
                                                 (long long)shard_size, reinterpret_cast<long long*>(d_out));
    CU(cudaGetLastError());
    ctx->launches++;
    if (st == ctx->stream) ctx->prev.valid = false;
    return DYNOED_OK;
}

int dynoed_set_profiling(dynoed_ctx* ctx, int on) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    if (on && !ctx->ev_k0) { CU(cudaEventCreate(&ctx->ev_k0)); CU(cudaEventCreate(&ctx->ev_k1)); }
    ctx->profiling = on != 0;
    ctx->kernel_ms_sum = 0.0; ctx->kernel_ms_count = 0; ctx->ev_pending = false;
    return DYNOED_OK;
}

int dynoed_kernel_time(dynoed_ctx* ctx, double* mean_ms, int64_t* count) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (ctx->ev_pending) {
        CU(cudaEventSynchronize(ctx->ev_k1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1));
        ctx->kernel_ms_sum += ms; ctx->kernel_ms_count++;
        ctx->ev_pending = false;
    }
    if (mean_ms) *mean_ms = ctx->kernel_ms_count ? ctx->kernel_ms_sum / ctx->kernel_ms_count : 0.0;
    if (count) *count = ctx->kernel_ms_count;
    return DYNOED_OK;
}

int dynoed_trajectory(dynoed_ctx* ctx, const double* designs, int64_t n, double* X) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!designs || !X || n <= 0) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    CU(cudaSetDevice(ctx->device));
    const auto& mi = ctx->model->info;
    const size_t row = (size_t)(ctx->n_intervals + 1) * (mi.nz + mi.nF);
    const int kd = ptr_kind(designs);
    if (ptr_kind(X) != kd) return fail(ctx, DYNOED_EINVAL, "designs/X must both be host or both be device pointers");
    double *dd = nullptr, *dx = nullptr, *dc = nullptr;
    auto cleanup = [&]() { cudaFree(dd); cudaFree(dx); cudaFree(dc); };
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
        return fail(ctx, DYNOED_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    ctx->call_flags = 0;
    CUT(cudaMalloc(&dc, sizeof(double) * n));
    int rc;
    if (kd == 1) {
        rc = eval_device(ctx, designs, n, dc, nullptr, nullptr, X);
        if (!rc) CUT(cudaStreamSynchronize(ctx->stream));
    } else {
        CUT(cudaMalloc(&dd, sizeof(double) * n * ctx->n_var));
        CUT(cudaMalloc(&dx, sizeof(double) * n * row));
        CUT(cudaMemcpyAsync(dd, designs, sizeof(double) * n * ctx->n_var, cudaMemcpyHostToDevice, ctx->stream));
        rc = eval_device(ctx, dd, n, dc, nullptr, nullptr, dx);
        if (!rc) {
            CUT(cudaMemcpyAsync(X, dx, sizeof(double) * n * row, cudaMemcpyDeviceToHost, ctx->stream));
            CUT(cudaStreamSynchronize(ctx->stream));
        }
    }
#undef CUT
    cleanup();
    return rc;
}

int dynoed_information_gain(dynoed_ctx* ctx, const double* designs, int64_t n, double* F, double* Pl, double* Pi) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!designs || n <= 0 || (!Pl && !Pi && !F)) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    CU(cudaSetDevice(ctx->device));
    const auto& mi = ctx->model->info;
    const int N = ctx->n_intervals;
    const size_t xrow = (size_t)(N + 1) * (mi.nz + mi.nF);
    const size_t prow = (size_t)(N + 1) * mi.n_obs * mi.np * mi.np;
    const int kd = ptr_kind(designs);
    if ((F && ptr_kind(F) != kd) || (Pl && ptr_kind(Pl) != kd) || (Pi && ptr_kind(Pi) != kd))
        return fail(ctx, DYNOED_EINVAL, "all pointers must be host or all device");
    double *dd = nullptr, *dx = nullptr, *dc = nullptr, *dP = nullptr, *dPi = nullptr, *dF = nullptr;
    int rc = DYNOED_OK;
    auto cleanup = [&]() { cudaFree(dd); cudaFree(dx); cudaFree(dc); cudaFree(dF); if (kd == 0) { cudaFree(dP); cudaFree(dPi); } };
#define CUI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
        return fail(ctx, DYNOED_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    CUI(cudaMalloc(&dx, sizeof(double) * n * xrow));
    CUI(cudaMalloc(&dc, sizeof(double) * n));
    CUI(cudaMalloc(&dF, sizeof(double) * n * mi.nF));
    const double* ddes = designs;
    if (kd == 0) {
        CUI(cudaMalloc(&dd, sizeof(double) * n * ctx->n_var));
        CUI(cudaMemcpyAsync(dd, designs, sizeof(double) * n * ctx->n_var, cudaMemcpyHostToDevice, ctx->stream));
        ddes = dd;
        if (Pl) CUI(cudaMalloc(&dP, sizeof(double) * n * prow));
        if (Pi) CUI(cudaMalloc(&dPi, sizeof(double) * n * prow));
    } else { dP = Pl; dPi = Pi; }
    ctx->call_flags = 0;
    rc = eval_device(ctx, ddes, n, dc, nullptr, dF, dx);     // trajectory at the grid points + F(t_f)
    if (rc) { cleanup(); return rc; }
    EvalParams P = ctx->base;
    P.designs = ddes; P.n = n;
    CUI(ctx->model->info_gain(P, dx, dP, dPi, ctx->stream));
    ctx->launches++;
    if (kd == 0) {
        if (F) CUI(cudaMemcpyAsync(F, dF, sizeof(double) * n * mi.nF, cudaMemcpyDeviceToHost, ctx->stream));
        if (Pl) CUI(cudaMemcpyAsync(Pl, dP, sizeof(double) * n * prow, cudaMemcpyDeviceToHost, ctx->stream));
        if (Pi) CUI(cudaMemcpyAsync(Pi, dPi, sizeof(double) * n * prow, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (F) {
        CUI(cudaMemcpyAsync(F, dF, sizeof(double) * n * mi.nF, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUI(cudaStreamSynchronize(ctx->stream));
#undef CUI
    cleanup();
    return DYNOED_OK;
}

// Hessians of the objective for n points: central differences of the exact gradient, all
// n (2 n_var + 1) perturbed designs evaluated as ordinary batches of the gradient kernel.
int dynoed_hessian_batch(dynoed_ctx* ctx, const double* designs, int64_t n, double rel_step, double* crit,
                         double* grad, double* hess) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!designs || !hess || n <= 0) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    if (!(rel_step >= 0.0) || rel_step > 0.1) return fail(ctx, DYNOED_EINVAL, "rel_step must lie in [0, 0.1] (0 = default)");
    CU(cudaSetDevice(ctx->device));
    const int kd = ptr_kind(designs);
    if (ptr_kind(hess) != kd || (crit && ptr_kind(crit) != kd) || (grad && ptr_kind(grad) != kd))
        return fail(ctx, DYNOED_EINVAL, "designs/crit/grad/hess must all be host or all be device pointers");
    const int nv = ctx->n_var;
    const int64_t R = 2 * (int64_t)nv + 1;
    const double rel = rel_step > 0.0 ? rel_step : 6.0554544523933395e-06;   // cbrt(2^-52)
    const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1, 131072 / R));
    auto& hs = ctx->hs;
    if (chunk > hs.cap) {
        CU(cudaDeviceSynchronize());
        cudaFree(hs.pert); cudaFree(hs.pcrit); cudaFree(hs.pgrad); cudaFree(hs.x); cudaFree(hs.c); cudaFree(hs.g); cudaFree(hs.H);
        hs = dynoed_ctx::HessScratch();
        CU(cudaMalloc(&hs.pert, sizeof(double) * chunk * R * nv));
        CU(cudaMalloc(&hs.pcrit, sizeof(double) * chunk * R));
        CU(cudaMalloc(&hs.pgrad, sizeof(double) * chunk * R * nv));
        CU(cudaMalloc(&hs.x, sizeof(double) * chunk * nv));
        CU(cudaMalloc(&hs.c, sizeof(double) * chunk));
        CU(cudaMalloc(&hs.g, sizeof(double) * chunk * nv));
        CU(cudaMalloc(&hs.H, sizeof(double) * chunk * nv * nv));
        hs.cap = chunk;
    }
    cudaStream_t st = ctx->stream;
    ctx->kidx_req = 1;
    ctx->call_flags = 0;
    ctx->prev.valid = false;
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t m = std::min(chunk, n - o);
        const double* xd = designs + o * nv;
        double *cd = crit ? crit + o : nullptr, *gd = grad ? grad + o * nv : nullptr, *Hd = hess + o * nv * nv;
        if (kd == 0) {
            CU(cudaMemcpyAsync(hs.x, xd, sizeof(double) * m * nv, cudaMemcpyHostToDevice, st));
            xd = hs.x; cd = crit ? hs.c : nullptr; gd = grad ? hs.g : nullptr; Hd = hs.H;
        }
        const int64_t tot_p = m * R * nv, tot_h = m * nv * (int64_t)nv;
        fd_perturb_kernel<<<(int)std::min<int64_t>((tot_p + 255) / 256, 4096), 256, 0, st>>>(xd, m, nv, rel, hs.pert);
        CU(cudaGetLastError());
        int rc = eval_device(ctx, hs.pert, m * R, hs.pcrit, hs.pgrad, nullptr, nullptr);
        if (rc) return rc;
        fd_hessian_kernel<<<(int)std::min<int64_t>((tot_h + 255) / 256, 4096), 256, 0, st>>>(hs.pert, hs.pcrit, hs.pgrad, m, nv, cd, gd, Hd);
        CU(cudaGetLastError());
        ctx->launches += 2;
        ctx->prev.valid = false;      // the scratch is reused by the next chunk: no overlap across it
        if (kd == 0) {
            if (crit) CU(cudaMemcpyAsync(crit + o, hs.c, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
            if (grad) CU(cudaMemcpyAsync(grad + o * nv, hs.g, sizeof(double) * m * nv, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(hess + o * nv * nv, hs.H, sizeof(double) * m * nv * nv, cudaMemcpyDeviceToHost, st));
        }
    }
    CU(cudaStreamSynchronize(st));
    return DYNOED_OK;
}

int dynoed_fim_basis(dynoed_ctx* ctx, const double* design, double* basis) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!design || !basis) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    if (ptr_kind(design) != 0 || ptr_kind(basis) != 0) return fail(ctx, DYNOED_EINVAL, "host pointers expected");
    const auto& mi = ctx->model->info;
    const int nv = ctx->n_var, nF = mi.nF;
    // one unit-weight design per weight column; every other weight zero, controls / ICs as given
    std::vector<int> cols;
    for (int i = 0; i < mi.n_obs; ++i)
        for (int k = 0; k < ctx->w_len[i]; ++k) cols.push_back(ctx->base.w_off[i] + k);
    std::vector<double> batch(cols.size() * (size_t)nv), crit(cols.size()), fim(cols.size() * (size_t)nF);
    for (size_t r = 0; r < cols.size(); ++r) {
        double* row = batch.data() + r * nv;
        memcpy(row, design, sizeof(double) * nv);
        for (int c : cols) row[c] = 0.0;
        row[cols[r]] = 1.0;
    }
    int rc = eval_common(ctx, batch.data(), (int64_t)cols.size(), crit.data(), nullptr, fim.data(), true, true);
    if (rc) return rc;
    memset(basis, 0, sizeof(double) * (size_t)nv * nF);
    for (size_t r = 0; r < cols.size(); ++r) memcpy(basis + (size_t)cols[r] * nF, fim.data() + r * nF, sizeof(double) * nF);
    return DYNOED_OK;
}

int dynoed_eval_fixed_controls(dynoed_ctx* ctx, const double* basis, const double* designs, int64_t n, double* crit,
                               double* grad, double* fim) {
    if (!ctx) return fail(nullptr, DYNOED_EINVAL, "ctx is NULL");
    if (!basis || !designs || !crit || n <= 0) return fail(ctx, DYNOED_EINVAL, "bad arguments");
    CU(cudaSetDevice(ctx->device));
    const auto& mi = ctx->model->info;
    const int nv = ctx->n_var, nF = mi.nF;
    const int kd = ptr_kind(designs);
    if (ptr_kind(crit) != kd || (grad && ptr_kind(grad) != kd) || (fim && ptr_kind(fim) != kd))
        return fail(ctx, DYNOED_EINVAL, "designs/crit/grad/fim must all be host or all be device pointers");
    if (mi.np > 6) return fail(ctx, DYNOED_EUNSUPPORTED, "np > 6");
    cudaStream_t st = ctx->stream;
    double *db = nullptr, *dd = nullptr, *dc = nullptr, *dg = nullptr, *df = nullptr;
    auto cleanup = [&]() { cudaFree(db); cudaFree(dd); cudaFree(dc); cudaFree(dg); cudaFree(df); };
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
        return fail(ctx, DYNOED_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    const double* bdev = basis;
    if (ptr_kind(basis) == 0) {
        CUF(cudaMalloc(&db, sizeof(double) * nv * nF));
        CUF(cudaMemcpyAsync(db, basis, sizeof(double) * nv * nF, cudaMemcpyHostToDevice, st));
        bdev = db;
    }
    const double* ddes = designs; double *dcr = crit, *dgr = grad, *dfi = fim;
    if (kd == 0) {
        CUF(cudaMalloc(&dd, sizeof(double) * n * nv)); CUF(cudaMalloc(&dc, sizeof(double) * n));
        if (grad) CUF(cudaMalloc(&dg, sizeof(double) * n * nv));
        if (fim) CUF(cudaMalloc(&df, sizeof(double) * n * nF));
        CUF(cudaMemcpyAsync(dd, designs, sizeof(double) * n * nv, cudaMemcpyHostToDevice, st));
        ddes = dd; dcr = dc; dgr = dg; dfi = df;
    }
    const int block = kFcTile;
    const size_t smem = 128 + sizeof(double) * ((size_t)kFcTile * nv + (size_t)nv * nF);
    if (smem > ctx->smem_optin) { cleanup(); return fail(ctx, DYNOED_EUNSUPPORTED, "design rows too long for the fixed-controls kernel"); }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ctx->smem_per_sm / (smem + 1024)));
    const int64_t ntiles = (n + kFcTile - 1) / kFcTile;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * per_sm);
    const int use_tma = (reinterpret_cast<uintptr_t>(ddes) % 16 == 0 && (!dgr || reinterpret_cast<uintptr_t>(dgr) % 16 == 0) &&
                         ((size_t)kFcTile * nv * 8) % 16 == 0) ? 1 : 0;
    switch (mi.np) {
#define DYNOED_FC(NPV) case NPV: { auto kern = fixed_controls_kernel<NPV>; \
        CUF(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, block, smem, st>>>(ddes, bdev, n, nv, ctx->base.tau_off, ctx->criterion, use_tma, dcr, dgr, dfi); } break;
        DYNOED_FC(1) DYNOED_FC(2) DYNOED_FC(3) DYNOED_FC(4) DYNOED_FC(5) DYNOED_FC(6)
#undef DYNOED_FC
    }
    CUF(cudaGetLastError());
    ctx->launches++;
    if (kd == 0) {
        CUF(cudaMemcpyAsync(crit, dc, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (grad) CUF(cudaMemcpyAsync(grad, dg, sizeof(double) * n * nv, cudaMemcpyDeviceToHost, st));
        if (fim) CUF(cudaMemcpyAsync(fim, df, sizeof(double) * n * nF, cudaMemcpyDeviceToHost, st));
        CUF(cudaStreamSynchronize(st));
    } else if (db) {
        CUF(cudaStreamSynchronize(st));
    }
#undef CUF
    cleanup();
    return DYNOED_OK;
}

int dynoed_criterion_batch(int device, int criterion, int np, const double* F, const double* tau, int64_t batch,
                           double* value, double* dcdF) {
    dynoed_ctx* ctx = nullptr;
    if (criterion < 0 || criterion >= kNumCriteria) return fail(nullptr, DYNOED_EINVAL, "criterion out of range");
    if (np < 1 || np > kMaxCritN) return fail(nullptr, DYNOED_EINVAL, "np must be 1..24");
    if (!F || !value || batch <= 0) return fail(nullptr, DYNOED_EINVAL, "bad arguments");
    CU(cudaSetDevice(device));
    const int nF = np * (np + 1) / 2;
    const int kd = ptr_kind(F);
    if (ptr_kind(value) != kd || (tau && ptr_kind(tau) != kd) || (dcdF && ptr_kind(dcdF) != kd))
        return fail(nullptr, DYNOED_EINVAL, "all pointers must be host or all device");
    const int block = 64;
    const int grid = (int)((batch + block - 1) / block);
    if (kd == 1) {
        criterion_batch_kernel<<<grid, block>>>(F, tau, np, batch, criterion, value, dcdF);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
        return DYNOED_OK;
    }
    double *dF = nullptr, *dt = nullptr, *dv = nullptr, *dL = nullptr;
    CU(cudaMalloc(&dF, sizeof(double) * batch * nF));
    CU(cudaMalloc(&dv, sizeof(double) * batch));
    CU(cudaMemcpy(dF, F, sizeof(double) * batch * nF, cudaMemcpyHostToDevice));
    if (tau) { CU(cudaMalloc(&dt, sizeof(double) * batch)); CU(cudaMemcpy(dt, tau, sizeof(double) * batch, cudaMemcpyHostToDevice)); }
    if (dcdF) CU(cudaMalloc(&dL, sizeof(double) * batch * nF));
    criterion_batch_kernel<<<grid, block>>>(dF, dt, np, batch, criterion, dv, dL);
    CU(cudaGetLastError());
    CU(cudaMemcpy(value, dv, sizeof(double) * batch, cudaMemcpyDeviceToHost));
    if (dcdF) CU(cudaMemcpy(dcdF, dL, sizeof(double) * batch * nF, cudaMemcpyDeviceToHost));
    cudaFree(dF); cudaFree(dt); cudaFree(dv); cudaFree(dL);
    return DYNOED_OK;
}

int dynoed_fp64_peak(int device, void* stream, double* tflops, double* ms_out) {
    dynoed_ctx* ctx = nullptr;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256, grid = prop.multiProcessorCount * 8, iters = 4096;
    double* out = nullptr;
    CU(cudaMalloc(&out, sizeof(double) * (size_t)grid * block));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0, st));
        fp64_peak_kernel<<<grid, block, 0, st>>>(out, iters, 1.0);
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    CU(cudaGetLastError());
    const double flops = 2.0 * 8 * 16 * (double)iters * grid * block;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return DYNOED_OK;
}

// instruction throughput of FP64 operand patterns, in Ginstr/s per GPU (thread-level instructions)
int dynoed_fp64_microbench(int device, int variant, double* ginstr_per_s) {
    dynoed_ctx* ctx = nullptr;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int block = 256, grid = prop.multiProcessorCount * 8, iters = 2048;
    double* out = nullptr;
    CU(cudaMalloc(&out, sizeof(double) * (size_t)grid * block));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, 0));
        switch (variant) {
            case 1: fp64_variant_kernel<1><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 2: fp64_variant_kernel<2><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 3: fp64_variant_kernel<3><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 4: fp64_variant_kernel<4><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 5: fp64_variant_kernel<5><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 6: fp64_variant_kernel<6><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 7: fp64_variant_kernel<7><<<grid, block>>>(out, iters, 1.0, 0.999999); break;
            case 8: dmma_probe_kernel<0><<<grid, block>>>(out, iters, 1.0); break;
            case 9: dmma_probe_kernel<1><<<grid, block>>>(out, iters, 1.0); break;
            default: fp64_peak_kernel<<<grid, block>>>(out, iters, 1.0); break;
        }
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    CU(cudaGetLastError());
    // variants 0..7: 8 chains x 16 FP64 instructions per iteration and thread; 8: 16 DMMA (each
    // 8x8x4 = 256 FMA per warp = 8 per thread); 9: 16 DMMA + 32 DFMA per iteration and thread
    double per_iter = 8.0 * 16;
    if (variant == 8) per_iter = 16.0 * 8;            // thread-level FMA equivalents
    if (variant == 9) per_iter = 16.0 * 8 + 32.0;
    if (ginstr_per_s) *ginstr_per_s = per_iter * (double)iters * grid * block / (best * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return DYNOED_OK;
}

int dynoed_kernel_stats_query(dynoed_ctx* ctx, int with_grad, dynoed_kernel_stats* o) {
    if (!ctx || !o) return fail(ctx, DYNOED_EINVAL, "ctx/out is NULL");
    const int g = with_grad == 2 ? 2 : with_grad ? 1 : 0;
    const auto& mi = ctx->model->info;
    memset(o, 0, sizeof(*o));
    o->regs_per_thread = ctx->fa[g].numRegs;
    o->static_smem = (int)ctx->fa[g].sharedSizeBytes;
    o->dynamic_smem = (int)smem_bytes(ctx, g);
    o->blocks_per_sm = ctx->bps[g];
    o->grid = ctx->sm_count * ctx->bps[g];
    o->block = ctx->tile;
    o->sm_count = ctx->sm_count;
    o->launches = ctx->launches;
    o->overlapped_launches = ctx->pdl_launches;
    o->small_batch_launches = ctx->dir_launches;
    o->small_batch_max = ctx->dir_ok ? (int32_t)(ctx->dir_max_lanes / ctx->dir_nl) : 0;
    o->time_parallel_max = ctx->tp_ok ? (int32_t)std::min<int64_t>(ctx->tp_max_n, o->small_batch_max) : 0;
    o->scratch_bytes = (int64_t)ctx->scratch_bytes;
    o->n_steps = ctx->n_steps;
    o->stages = ctx->stages;
    // algorithmic flops of the straight-line code the emitter produced plus the template's own
    // arithmetic (FMA = 2, mul/add = 1):
    //   forward step : S x (primal_z + obs) + stage states (2 NZ per non-zero a_ij) + z accumulation
    //                  (2 NZ per stage) + unweighted quadrature accumulation (2 NOBS NF per stage, plus
    //                  NOBS NP scalings for every stage with b_s != b_0)
    //   per interval : P = (h b_0) acc, F += w P (3 NOBS NF), consts
    //   backward step: (S-1) primal_z + S adjoint + stage states + kb' (2 NZ per a_ij each) +
    //                  lambda / u-bar accumulation (2 per entry and stage)
    //   per interval : cwL = 2 w L (NOBS (1 + NF)), d/dw = <Lambda, P> (2 NOBS NF), consts
    const int S = ctx->stages;
    int nza = 0, nscaled = 0;
    for (int i = 0; i < S; ++i) {
        const double bi = ctx->method == DYNOED_RK4 ? TabRK4::b(i) : TabTsit5::b(i);
        const double b0 = ctx->method == DYNOED_RK4 ? TabRK4::b(0) : TabTsit5::b(0);
        nscaled += (bi / b0 != 1.0);
        for (int j = 0; j < i; ++j)
            nza += (ctx->method == DYNOED_RK4 ? TabRK4::a(i, j) : TabTsit5::a(i, j)) != 0.0;
    }
    const int nq = mi.n_obs * mi.nF;
    // stages with equal weights share an accumulator set (QuadClasses): no scalings, one more fold per set
    const int nsets = ctx->method == DYNOED_RK4 ? QuadClasses<TabRK4>::N : QuadClasses<TabTsit5>::N;
    if (nsets > 1) nscaled = 0;
    const double fold = 2.0 * nq * (nsets - 1);
    const double fwd = (double)S * mi.flops_primal + 2.0 * mi.nz * nza + 2.0 * mi.nz * S + 2.0 * nq * S +
                       (double)nscaled * mi.n_obs * mi.np + (ctx->ucoef ? 0.0 : 2.0 * nq + fold);
    const double fwd_iv = (ctx->ucoef ? 3.0 * nq + fold : 2.0 * nq) + mi.flops_consts;
    const double bwd = (double)(S - 1) * mi.flops_primal_z + (double)S * mi.flops_adjoint + 4.0 * mi.nz * nza +
                       2.0 * S * (mi.nz + mi.n_ctrl);
    const double bwd_iv = (double)mi.n_obs * (1 + mi.nF) + 2.0 * nq + mi.flops_consts;
    const double epi = 10.0 * mi.nF;
    if (ctx->rosenbrock) {
        // Rosenbrock kernels (config 4), algorithmic flops per design, FMA = 2:
        //   step   : f_x (ros_jac) + nx for W = c0 I - f_x + pivoted LU (1 + (n-1-k) + 2 (n-1-k)^2 per pivot)
        //   stage  : stage state (2 NZ per a_ij) and g = ros_f unless the tableau reuses the previous stage
        //            state; right-hand side (2 (NZ + NQ) per c_ij); 1 + np triangular solves (2 nx (nx-1) + nx
        //            each) + ros_ling + nx np; quadrature rows ros_linp + 2 NQ
        //   update : 2 (NZ + NQ) per stage with m_s != 0;   interval: F += w P (2 NQ)
        //   every gradient DIRECTION adds the tangent of all of that except the factorisation: twice the
        //   function / combination flops (dual-number product rule) + its own 1 + np solves with the value
        //   factors + ros_hvp + nx per solve; a pass carries NDIR directions and recomputes the values.
        const int nx = mi.nx, np_ = mi.np, nz = mi.nz, nqq = mi.n_obs * mi.nF;
        int Sr = 0, na = 0, nc = 0, nm = 0, nf = 0;
        auto tab = [&](auto RT) {
            using R = decltype(RT);
            Sr = R::S;
            for (int i = 0; i < R::S; ++i) {
                nm += R::m(i) != 0.0;
                nf += !R::same_stage_state(i);
                for (int j = 0; j < i; ++j) { na += (!R::same_stage_state(i)) && R::a(i, j) != 0.0; nc += R::c(i, j) != 0.0; }
            }
        };
        if (ctx->method == DYNOED_ROS2) tab(TabROS2{}); else if (ctx->method == DYNOED_ROS3P) tab(TabROS3P{}); else tab(TabRODAS3{});
        double lu = 0.0;
        for (int k = 0; k < nx; ++k) lu += 1.0 + (nx - 1 - k) + 2.0 * (nx - 1 - k) * (nx - 1 - k);
        const double solve = 2.0 * nx * (nx - 1) + nx;
        const double fun = (double)nf * mi.flops_ros_f + (double)Sr * (ctx->model->flops_ros_ling + ctx->model->flops_ros_linp);
        const double comb = 2.0 * nz * na + (2.0 * (nz + nqq) + 1.0) * nc + (double)Sr * (nx * np_ + 2.0 * nqq) + 2.0 * (nz + nqq) * nm;
        const double solves = (double)Sr * (1 + np_) * solve;
        const double val_step = ctx->model->flops_ros_jac + nx + lu + fun + comb + solves;
        const double dir_step = 2.0 * (fun + comb) + solves + (double)Sr * (1 + np_) * (ctx->model->flops_ros_hvp + nx);
        const double val = val_step * ctx->n_steps + (double)ctx->n_intervals * (2.0 * nqq + mi.flops_consts) + epi;
        const bool cgrad = g == 1 && mi.n_ctrl > 0;
        const int ndirs = mi.n_ic + (cgrad ? ctx->base.n_ctrl_values : 0);
        const int cap = Sr >= 4 ? 1 : DYNOED_ROS_NDIR;   // ros_ndir()
        const int per_pass = cgrad ? cap : std::max(1, std::min(cap, (int)mi.n_ic));
        const int passes = ndirs > 0 ? (ndirs + per_pass - 1) / per_pass : 1;
        o->flops_per_design_eval = val;
        // d/dw, d/dtau from the stored quadratures: 2 NQ per interval
        o->flops_per_design_grad = passes * val + ndirs * (dir_step * ctx->n_steps + 4.0 * nqq * ctx->n_intervals) +
                                   2.0 * nqq * ctx->n_intervals;
        if (g == 1 && ctx->ic2) {
            // ros_ic_kernel: ONE value pass + the second-order columns.  Per stage: the emitted right-hand
            // sides / couplings of all NH columns and of the dP rows, NH solves with the value factors
            // (+ nx for r + b); per step: column stage states, c_ij and m_s combinations (2 nx each) and the
            // stage recursion of the NQ NIC quadrature-derivative rows.
            int na_all = 0;
            auto cnt = [&](auto RT) { using R = decltype(RT); for (int i = 0; i < R::S; ++i) for (int j = 0; j < i; ++j) na_all += R::a(i, j) != 0.0; };
            if (ctx->method == DYNOED_ROS2) cnt(TabROS2{}); else if (ctx->method == DYNOED_ROS3P) cnt(TabROS3P{}); else cnt(TabRODAS3{});
            const int nh = ctx->model->nh, nq2 = nqq * mi.n_ic;
            const double h_step = (double)Sr * (ctx->model->flops_ic2_col + ctx->model->flops_ic2_q + nh * (solve + nx)) +
                                  2.0 * nh * nx * (na_all + nc + nm) + (double)nq2 * (2.0 * nc + Sr + 2.0 * nm);
            o->flops_per_design_grad = val + h_step * ctx->n_steps +
                                       (double)ctx->n_intervals * (2.0 * nq2 + 2.0 * nqq) + 2.0 * mi.nF * mi.n_ic;
        }
        o->bytes_per_design = 8.0 * (ctx->n_var + 1 + (with_grad ? ctx->n_var : 0) + mi.nF);
        return DYNOED_OK;
    }
    o->flops_per_design_eval = fwd * ctx->n_steps + (double)ctx->n_intervals * fwd_iv + epi;
    o->flops_per_design_grad = (fwd + bwd) * ctx->n_steps + (double)ctx->n_intervals * (fwd_iv + bwd_iv) + epi;
    if (g == 2) {   // forward-only gradient: unweighted per-observation quadratures, F += w P per interval,
                    // d/dw = <Lambda, P> in the epilogue
        if (mi.n_ic == 0) {   // oed_wgrad_kernel: the objective kernel's sweep + <Lambda, P_ij> per interval
            o->flops_per_design_grad = o->flops_per_design_eval + (double)ctx->n_intervals * 2.0 * nq + mi.np;
        } else {
            const double f2 = (double)S * mi.flops_ros_f + 2.0 * mi.nz * nza + 2.0 * (mi.nz + nq) * S + nza + S;
            o->flops_per_design_grad = f2 * ctx->n_steps + (double)ctx->n_intervals * (mi.flops_consts + 4.0 * nq) + epi;
        }
    }
    o->bytes_per_design = 8.0 * (ctx->n_var + 1 + (with_grad ? ctx->n_var : 0) + mi.nF);
    return DYNOED_OK;
}

int dynoed_host_alloc(void** p, size_t bytes) {
    dynoed_ctx* ctx = nullptr;
    if (!p) return fail(nullptr, DYNOED_EINVAL, "p is NULL");
    CU(cudaMallocHost(p, bytes));
    return DYNOED_OK;
}

int dynoed_host_free(void* p) {
    dynoed_ctx* ctx = nullptr;
    CU(cudaFreeHost(p));
    return DYNOED_OK;
}

}  // extern "C"
