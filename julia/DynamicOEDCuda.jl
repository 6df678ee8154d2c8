# DynamicOEDCuda.jl — thin `ccall` shim that plugs libdynoed_cuda.so into DynamicOED.jl.
#
# STATUS: written against DynamicOED.jl's sources (src/problem.jl:96-161, src/discretize.jl:32-173),
# NOT executed — Julia is not available in the build image.  The same ABI is exercised from Python
# (dynamicoed.jl_b200/runtime.py, tests/test_abi.py, tests/test_gpu_parity.py).
#
# What it does: `OptimizationProblem(prob::OEDProblem, ::CudaBackend, u0, p; kwargs...)` builds
# everything exactly as src/problem.jl:96-161 does — bounds, integrality, constraints, `syms` — but
# hands Optimization.jl an `OptimizationFunction(f; grad = g!, cons = ...)` whose objective and
# gradient are evaluated by the GPU library.  No CUDA.jl codegen, no CPU fallback: if the library
# or a device is missing the constructor throws.
#
#   using DynamicOED, DynamicOEDCuda, Optimization, OptimizationMOI, Ipopt
#   prob = OEDProblem(structural_simplify(OEDSystem(sys)), FisherDCriterion())
#   optp = OptimizationProblem(prob, CudaBackend(model = "lv_fishing", alg = :rk4, substeps = 4))
#   res  = solve(optp, Ipopt.Optimizer())      # exact Hessian from the GPU; "limited-memory" works too
module DynamicOEDCuda

using DynamicOED
using DynamicOED: OEDProblem, Timegrid, generate_initial_variables, generate_variable_bounds,
                  get_initial_conditions, get_control_parameters, get_measurement_function,
                  AbstractInformationCriterion
using ComponentArrays
using ModelingToolkit
using Optimization
using SciMLBase
using LinearAlgebra: logdet

export CudaBackend, DynoedContext, eval_batch!, hessian!, eval_batch_async!, wait_call, argmin_design, dynoed_library,
       pinned_matrix

const LIB = Ref{String}(get(ENV, "DYNOED_CUDA_LIB", "libdynoed_cuda.so"))
dynoed_library() = LIB[]
dynoed_library(path::AbstractString) = (LIB[] = String(path))

# ---- mirrors of include/dynoed.h -------------------------------------------------------------
const DYNOED_OK = Cint(0)
const METHODS = Dict(:rk4 => Cint(0), :tsit5 => Cint(1), :ros2 => Cint(2), :ros3p => Cint(3), :rodas3 => Cint(4))

criterion_id(::FisherACriterion) = Cint(0)
criterion_id(::FisherDCriterion) = Cint(1)
criterion_id(::FisherECriterion) = Cint(2)
criterion_id(::ACriterion) = Cint(3)
criterion_id(::DCriterion) = Cint(4)
criterion_id(::ECriterion) = Cint(5)
# extension (not in DynamicOED.jl): -log det(F + tau I), DYNOED_LOG_D
struct LogDCriterion <: DynamicOED.AbstractInformationCriterion end
(c::LogDCriterion)(F::AbstractArray{T, 2}) where {T} = -logdet(F)
criterion_id(::LogDCriterion) = Cint(6)

# struct dynoed_problem_desc (include/dynoed.h) — field order and types must match exactly
struct ProblemDesc
    struct_size::UInt32
    model::Cstring
    criterion::Int32
    method::Int32
    n_intervals::Int32
    tgrid::Ptr{Float64}
    substeps::Int32
    edge_count::Ptr{Int32}
    step_edges::Ptr{Float64}
    indicators::Ptr{Int32}
    ctrl_offsets::Ptr{Int32}
    w_offsets::Ptr{Int32}
    ic_offset::Int32
    tau_offset::Int32
    n_var::Int32
    theta::Ptr{Float64}
    x0::Ptr{Float64}
    G0::Ptr{Float64}
    tile::Int32
    max_batch_hint::Int32
end

"""
    CudaBackend(; model, alg = :rk4, substeps = 4, device = 0)

Stands where the `AD::AbstractADType` argument of `OptimizationProblem(::OEDProblem, AD, ...)`
stands (src/problem.jl:96-97).  `model` names a model compiled into the library (emitted by
`dynamicoed.jl_b200/emit.py`; `dynoed_model_name(i)` lists them).
"""
Base.@kwdef struct CudaBackend <: Optimization.ADTypes.AbstractADType
    model::String
    alg::Symbol = :rk4
    substeps::Int = 4
    device::Int = 0
end

mutable struct DynoedContext
    handle::Ptr{Cvoid}
    n_var::Int
    nF::Int
    keep::Vector{Any}          # arrays the descriptor points to (kept alive until create returns)
end

function last_error(h::Ptr{Cvoid})
    p = ccall((:dynoed_last_error, LIB[]), Cstring, (Ptr{Cvoid},), h)
    p == C_NULL ? "?" : unsafe_string(p)
end

check(rc::Cint, h = C_NULL) = rc == DYNOED_OK || error("dynoed error $rc: $(last_error(h))")

# ---- design-vector layout: offsets of the blocks inside the flat ComponentVector --------------
# (src/discretize.jl:110-137; keys sorted: controls | initial_conditions | measurements | regularization)
function layout_offsets(prob::OEDProblem)
    u0 = generate_initial_variables(prob.system, prob.timegrid)
    flat = collect(1:length(u0))
    idx = ComponentVector(flat, getaxes(u0))
    ctrl = Int32[first(getproperty(idx.controls, Symbol(c))) - 1 for c in get_control_parameters(prob.system)]
    w = Int32[first(getproperty(idx.measurements, Symbol(wi))) - 1 for wi in get_measurement_function(prob.system)]
    ics = get_initial_conditions(prob.system)
    ic = isempty(ics) ? Int32(0) : Int32(first(getproperty(idx.initial_conditions, Symbol(first(ics)))) - 1)
    tau = Int32(first(idx.regularization) - 1)
    ctrl, w, ic, tau, length(u0)
end

function DynoedContext(prob::OEDProblem, backend::CudaBackend)
    tg = prob.timegrid
    N = length(tg.timespans)
    tgrid = Float64[first.(tg.timespans); last(last(tg.timespans))]
    # Timegrid.indicators is [n_var x N], 1-based, rows ordered (controls..., w...) (src/discretize.jl:64-83)
    ind = Int32.(permutedims(tg.indicators .- 1))          # -> column-major [N x n_var] == row-major [n_var][N]
    ctrl, w, ic, tau, n_var = layout_offsets(prob)
    isempty(ctrl) && (ctrl = Int32[0])
    keep = Any[tgrid, ind, ctrl, w]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep backend begin
        desc = ProblemDesc(UInt32(sizeof(ProblemDesc)), Base.unsafe_convert(Cstring, backend.model),
            criterion_id(prob.objective), METHODS[backend.alg], Int32(N), pointer(tgrid),
            Int32(backend.substeps), C_NULL, C_NULL, pointer(ind), pointer(ctrl), pointer(w),
            ic, tau, Int32(n_var), C_NULL, C_NULL, C_NULL, Int32(0), Int32(0))
        rc = ccall((:dynoed_create, LIB[]), Cint, (Ref{ProblemDesc}, Cint, Ref{Ptr{Cvoid}}),
            Ref(desc), Cint(backend.device), h)
        check(rc)
    end
    nF = let f = count(DynamicOED.is_fisher_state, states(prob.system)); f end
    ctx = DynoedContext(h[], n_var, nF, keep)
    finalizer(c -> (c.handle != C_NULL && ccall((:dynoed_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), ctx)
    ctx
end

"""
    eval_batch!(ctx, designs, crit; grad = nothing, fim = nothing)

`designs` is `n_var x n` (column-major Julia matrix = the library's row-major `[n][n_var]`);
`crit` length n; `grad` like `designs`; `fim` is `nF x n`.  Host arrays; blocking.
Replaces the objective closure + ForwardDiff gradient of src/problem.jl:112-121, :154 for a batch.
"""
function eval_batch!(ctx::DynoedContext, designs::Matrix{Float64}, crit::Vector{Float64};
        grad::Union{Nothing, Matrix{Float64}} = nothing, fim::Union{Nothing, Matrix{Float64}} = nothing)
    n = size(designs, 2)
    @assert size(designs, 1) == ctx.n_var && length(crit) == n
    rc = ccall((:dynoed_eval_batch, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt32),
        ctx.handle, designs, n, crit, grad === nothing ? C_NULL : pointer(grad),
        fim === nothing ? C_NULL : pointer(fim), UInt32(0))
    check(rc, ctx.handle)
    crit
end

"""
    hessian!(ctx, H, u; crit = nothing, grad = nothing, rel_step = 0.0) -> H

Hessian of the objective at `u` (`n_var` vector, or an `n_var x n` matrix with `H` of size
`n_var x n_var x n`): central differences of the exact gradient, the `2 n_var + 1` perturbed designs
of every point evaluated as one batch of the gradient kernel (`dynoed_hessian_batch`).
"""
function hessian!(ctx::DynoedContext, H::Array{Float64}, u::Array{Float64};
        crit::Union{Nothing, Vector{Float64}} = nothing, grad::Union{Nothing, Array{Float64}} = nothing,
        rel_step::Float64 = 0.0)
    n = size(u, 2)
    @assert size(u, 1) == ctx.n_var && length(H) == ctx.n_var^2 * n
    rc = ccall((:dynoed_hessian_batch, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.handle, u, n, rel_step, crit === nothing ? C_NULL : pointer(crit),
        grad === nothing ? C_NULL : pointer(grad), H)
    check(rc, ctx.handle)
    H
end

"""
    pinned_matrix(rows, cols) -> Matrix{Float64}

Page-locked host memory from `dynoed_host_alloc`, wrapped as a Julia matrix (freed by a finalizer).
Ordinary Julia arrays are pageable: the library accepts them, but large batches then move at a few
GB/s instead of the PCIe rate (1 M LV designs: 71 ms pageable vs 15 ms pinned on B200).  Use this
for the `designs` / `crit` / `grad` buffers of batched sweeps.
"""
function pinned_matrix(rows::Integer, cols::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dynoed_host_alloc, LIB[]), Cint, (Ref{Ptr{Cvoid}}, Csize_t), p, rows * cols * sizeof(Float64)))
    A = unsafe_wrap(Array, Ptr{Float64}(p[]), (Int(rows), Int(cols)); own = false)
    finalizer(_ -> ccall((:dynoed_host_free, LIB[]), Cint, (Ptr{Cvoid},), p[]), A)
    A
end

"""
    id = eval_batch_async!(ctx, designs, crit; grad, fim);  wait_call(ctx, id)

Pipelined host-buffer evaluation for ONE Julia thread: returns at once; `wait_call` blocks until the
results are in `crit` / `grad` / `fim`.  Keep two calls in flight on alternating `pinned_matrix`
buffers and one call's device-to-host copy overlaps the next call's host-to-device copy
(1.01 vs 1.18 ms per 65,536-design step on B200).
"""
function eval_batch_async!(ctx::DynoedContext, designs::Matrix{Float64}, crit::AbstractVector{Float64};
        grad::Union{Nothing, Matrix{Float64}} = nothing, fim::Union{Nothing, Matrix{Float64}} = nothing)
    id = Ref{Int64}(0)
    rc = ccall((:dynoed_eval_batch_host_async, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt32, Ref{Int64}),
        ctx.handle, designs, size(designs, 2), crit, grad === nothing ? C_NULL : pointer(grad),
        fim === nothing ? C_NULL : pointer(fim), UInt32(0), id)
    check(rc, ctx.handle)
    id[]
end
wait_call(ctx::DynoedContext, id::Integer) =
    check(ccall((:dynoed_wait, LIB[]), Cint, (Ptr{Cvoid}, Int64), ctx.handle, id), ctx.handle)

function argmin_design(ctx::DynoedContext)
    i = Ref{Int64}(0); v = Ref{Float64}(0.0)
    check(ccall((:dynoed_argmin, LIB[]), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Float64}), ctx.handle, i, v), ctx.handle)
    Int(i[]) + 1, v[]
end

# ---- the drop-in: same signature as src/problem.jl:96-101 with CudaBackend in the AD slot ------
function Optimization.OptimizationProblem(prob::OEDProblem, backend::CudaBackend,
        u0::ComponentVector = DynamicOED.get_initial_variables(prob), p = SciMLBase.NullParameters();
        integer_constraints::Bool = false, constraints = nothing,
        variable_type::Type{T} = Float64, kwargs...) where {T}
    T === Float64 || error("the CUDA path computes in Float64 only (no dual numbers cross the C ABI)")
    u0 = Float64.(u0)
    ctx = DynoedContext(prob, backend)
    buf = zeros(ctx.n_var, 1); c1 = zeros(1); g1 = zeros(ctx.n_var, 1)

    # Ipopt asks for f(u) and grad f(u) at the same point in separate callbacks: one launch serves
    # both (the gradient kernel also returns the objective); `last` is a one-entry cache.
    last = fill(NaN, ctx.n_var)
    both! = u -> begin
        if !(last == u)                       # NaN-initialised => first call always evaluates
            buf[:, 1] .= u
            eval_batch!(ctx, buf, c1; grad = g1)
            last .= u
        end
        nothing
    end
    objective = (u, _) -> (both!(u); c1[1])
    gradient! = (G, u, _) -> (both!(u); G .= view(g1, :, 1); nothing)

    if isa(constraints, ConstraintsSystem)          # src/problem.jl:123-128
        (_, cons), cons_lb, cons_ub = ModelingToolkit.generate_function(constraints, expression = Val{false})
    else
        cons = cons_lb = cons_ub = nothing
    end
    lb = Float64.(generate_variable_bounds(prob.system, prob.timegrid, true))     # :131-132
    ub = Float64.(generate_variable_bounds(prob.system, prob.timegrid, false))
    @assert all(lb .<= u0 .<= ub) "The initial variables are not within the bounds. Please check the input!"
    integrality = Bool.(u0 * 0)                                                     # :137-146
    if integer_constraints
        integrality.measurements .= true
        for c in get_control_parameters(prob.system)
            getproperty(integrality.controls, Symbol(c)) .= ModelingToolkit.SymbolicUtils.symtype(c) <: Int
        end
    end
    syms = Symbol.(vcat(get_initial_conditions(prob.system), get_control_parameters(prob.system),
        get_measurement_function(prob.system)))
    # Objective, gradient AND Hessian come from the GPU, so `solve(opt_prob, Ipopt.Optimizer())` works
    # with Ipopt's default exact Hessian as in the reference's tests (test/references/1D.jl:30) as well
    # as with hessian_approximation = "limited-memory" (docs/src/examples/1D.md:89).  The AD type only
    # serves what is NOT supplied: Jacobian / Hessians of the pure-Julia constraint function `cons`.
    hbuf = Matrix{Float64}(undef, ctx.n_var, ctx.n_var)
    hessian_cb! = (H, u, _) -> begin
        buf[:, 1] .= u
        hessian!(ctx, hbuf, buf; crit = c1, grad = g1)     # refreshes the f / grad cache too
        last .= u
        H .= hbuf
        nothing
    end
    opt_f = OptimizationFunction(objective, Optimization.AutoForwardDiff(); grad = gradient!,
        hess = hessian_cb!, syms = syms, cons = cons)
    OptimizationProblem(opt_f, u0, p; lb = lb, ub = ub, int = integrality, lcons = cons_lb, ucons = cons_ub)
end

end # module
