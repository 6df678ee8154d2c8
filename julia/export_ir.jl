# export_ir.jl — writes the model interchange file ("dynoed-ir/1") that
# `python -m dynamicoed.jl_b200.ir build model.json` turns into a compiled kernel library.
#
# STATUS: written against ModelingToolkit ~8.72 / Symbolics as pinned by the reference's
# Project.toml, NOT executed (no Julia in the build image).  The format itself is pinned by
# tests/test_ir.py (Python round trip of every built-in model).
#
# The file describes the UN-augmented system exactly as the user declares it for DynamicOED.jl
# (e.g. /root/reference/test/references/LotkaVolterra.jl:10-36) AND, in its "augmented" section, the
# explicit right-hand sides of the reference's own `structural_simplify(OEDSystem(sys))`
# (`full_equations`: MTK's tearing result with the observed equations substituted back).  The Python
# host derives the structure the kernel emitter needs (Jacobian blocks, adjoint) from the user system
# and REFUSES to emit a kernel unless its augmented right-hand sides equal the exported ones
# (dynamicoed.jl_b200/ir.py::check_augmented): the MTK-simplified system is the authority for the
# equations that reach the kernel template.
#
#   include("export_ir.jl")
#   export_ir(lotka_volterra, "lv.json")                       # the plain ODESystem, before OEDSystem()
#   run(`python -m dynamicoed.jl_b200.ir build lv.json`)       # -> {"library": ..., "model": ..., "augmented_equations_checked": 9}
#   export_ir(sys, path; augmented = false)                    # user system only (no cross-check)
module DynoedIR

using ModelingToolkit
using ModelingToolkit: getdefault, hasdefault, istunable, isinput, getbounds, hasbounds
using Symbolics
using DynamicOED: is_measured, get_measurement_rate, OEDSystem, is_initial_condition, get_initial_condition_id,
                  is_measurement_function
using LinearAlgebra: triu

export export_ir

json_str(s) = "\"" * replace(String(s), "\\" => "\\\\", "\"" => "\\\"") * "\""
json_num(x::Nothing) = "null"
json_num(x::Bool) = x ? "true" : "false"
json_num(x::Real) = isfinite(x) ? repr(Float64(x)) : "null"
json_num(x::Integer) = string(x)

# infix text of an expression with every state/parameter replaced by an identifier-safe alias
function expr_text(ex, alias::Dict)
    ex = Symbolics.unwrap(ex)
    s = string(Symbolics.substitute(ex, alias))
    replace(s, "^" => "**")
end

# Canonical identifiers of the augmented system, from POSITIONS the reference fixes
# (/root/reference/src/augment.jl:233-243): states(OEDSystem(sys)) = [x; vec(G); F[triu]; vec(Q)],
# parameters = [union(p, p_tunable); w].  No name parsing.
function augmented_alias(sys, oed)
    x, p = states(sys), parameters(sys)
    nx = length(x)
    ps = parameters(oed)
    ic = filter(is_initial_condition, ps)
    w = filter(is_measurement_function, ps)
    np_ = count(v -> ModelingToolkit.istunable(v, false) && !ModelingToolkit.isinput(v), p) + length(ic)
    st = states(oed)
    alias = Dict{Any, Any}()
    sym(s) = Symbolics.unwrap(Symbolics.variable(Symbol(s)))
    for k in 1:nx
        alias[Symbolics.unwrap(st[k])] = sym("x$(k-1)")
    end
    for j in 1:np_, i in 1:nx                              # vec(G): column-major
        alias[Symbolics.unwrap(st[nx + (j - 1) * nx + i])] = sym("G_$(i)_$(j)")
    end
    k = nx + nx * np_
    for j in 1:np_, i in 1:j                               # F[triu(trues(np, np))]: column-major upper triangle
        k += 1
        alias[Symbolics.unwrap(st[k])] = sym("F_$(i)_$(j)")
    end
    for (k, v) in enumerate(p)
        alias[Symbolics.unwrap(v)] = sym("p$(k-1)")
    end
    for v in ic
        alias[Symbolics.unwrap(v)] = sym("ic$(get_initial_condition_id(v))")
    end
    for (i, v) in enumerate(w)
        alias[Symbolics.unwrap(v)] = sym("w$(i)")
    end
    alias
end

# "augmented": one entry per differential state of the reference's simplified augmented system
function augmented_section(sys)
    oed = OEDSystem(sys)
    alias = augmented_alias(sys, oed)
    simp = structural_simplify(oed)
    io = IOBuffer()
    print(io, "{\"equations\": [")
    first = true
    for eq in ModelingToolkit.full_equations(simp)
        lhs = Symbolics.unwrap(eq.lhs)
        (Symbolics.istree(lhs) && Symbolics.operation(lhs) isa Differential) || continue
        v = only(Symbolics.arguments(lhs))
        haskey(alias, v) || continue                       # Q rows are observed, not states
        first || print(io, ", ")
        first = false
        print(io, "{\"state\": ", json_str(string(alias[v])), ", \"rhs\": ", json_str(expr_text(eq.rhs, alias)), "}")
    end
    print(io, "]}")
    String(take!(io))
end

function export_ir(sys::ModelingToolkit.AbstractODESystem, path::AbstractString;
        tspan = ModelingToolkit.get_tspan(sys), augmented::Bool = true)
    t = ModelingToolkit.get_iv(sys)
    sts = states(sys)
    ps = parameters(sys)
    alias = Dict{Any, Any}()
    for (k, v) in enumerate(sts)
        alias[Symbolics.unwrap(v)] = Symbolics.unwrap(Symbolics.variable(Symbol("x", k - 1)))
    end
    for (k, v) in enumerate(ps)
        alias[Symbolics.unwrap(v)] = Symbolics.unwrap(Symbolics.variable(Symbol("p", k - 1)))
    end
    bounds(v) = hasbounds(v) ? getbounds(v) : (nothing, nothing)
    rate(v) = is_measured(v) ? get_measurement_rate(v) : nothing
    io = IOBuffer()
    print(io, "{\"format\": \"dynoed-ir/1\", \"name\": ", json_str(nameof(sys)), ", \"iv\": \"t\", ")
    print(io, "\"tspan\": ", tspan === nothing ? "null" : "[$(Float64(tspan[1])), $(Float64(tspan[2]))]", ",\n")
    print(io, "\"states\": [")
    for (k, v) in enumerate(sts)
        lo, hi = bounds(v)
        k > 1 && print(io, ", ")
        print(io, "{\"id\": \"x$(k-1)\", \"name\": ", json_str(string(Symbolics.tosymbol(v; escape = false))),
            ", \"default\": ", hasdefault(v) ? json_num(getdefault(v)) : "null",
            ", \"tunable\": ", json_num(istunable(v, false)),
            ", \"bounds\": [", json_num(lo), ", ", json_num(hi), "]}")
    end
    print(io, "],\n\"parameters\": [")
    for (k, v) in enumerate(ps)
        lo, hi = bounds(v)
        r = rate(v)
        k > 1 && print(io, ", ")
        print(io, "{\"id\": \"p$(k-1)\", \"name\": ", json_str(string(v)),
            ", \"default\": ", hasdefault(v) ? json_num(getdefault(v)) : "null",
            ", \"tunable\": ", json_num(istunable(v, false)), ", \"input\": ", json_num(isinput(v)),
            ", \"bounds\": [", json_num(lo), ", ", json_num(hi), "]",
            ", \"measurement_rate\": ", json_num(r), ", \"rate_is_count\": ", json_num(r isa Integer),
            ", \"integer\": ", json_num(Symbolics.symtype(Symbolics.unwrap(v)) <: Integer), "}")
    end
    print(io, "],\n\"equations\": [")
    for (k, eq) in enumerate(equations(sys))
        k > 1 && print(io, ", ")
        lhs = Symbolics.unwrap(eq.lhs)
        if Symbolics.istree(lhs) && Symbolics.operation(lhs) isa Differential
            x = only(Symbolics.arguments(lhs))
            print(io, "{\"lhs\": \"D(", string(alias[x]), ")\", \"rhs\": ", json_str(expr_text(eq.rhs, alias)), "}")
        else
            print(io, "{\"lhs\": ", json_str(expr_text(eq.lhs, alias)), ", \"rhs\": ", json_str(expr_text(eq.rhs, alias)), "}")
        end
    end
    print(io, "],\n\"observed\": [")
    for (k, eq) in enumerate(observed(sys))
        k > 1 && print(io, ", ")
        r = get_measurement_rate(eq.lhs)
        print(io, "{\"name\": ", json_str(string(Symbolics.tosymbol(eq.lhs; escape = false))),
            ", \"rhs\": ", json_str(expr_text(eq.rhs, alias)),
            ", \"measurement_rate\": ", json_num(r), ", \"rate_is_count\": ", json_num(r isa Integer), "}")
    end
    print(io, "]")
    augmented && print(io, ",\n\"augmented\": ", augmented_section(sys))
    print(io, "}\n")
    write(path, String(take!(io)))
    path
end

end # module
