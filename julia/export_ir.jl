# export_ir.jl — writes the model interchange file ("dynoed-ir/1") that
# `python -m dynamicoed.jl_b200.ir build model.json` turns into a compiled kernel library.
#
# STATUS: written against ModelingToolkit ~8.72 / Symbolics as pinned by the reference's
# Project.toml, NOT executed (no Julia in the build image).  The format itself is pinned by
# tests/test_ir.py (Python round trip of every built-in model).
#
# The file describes the UN-augmented system exactly as the user declares it for DynamicOED.jl
# (e.g. /root/reference/test/references/LotkaVolterra.jl:10-36): the augmentation
# (`OEDSystem`, src/augment.jl:128-244) and `structural_simplify` are re-done by the Python host
# with the same equations, so nothing MTK-internal has to cross the boundary.
#
#   include("export_ir.jl")
#   export_ir(lotka_volterra, "lv.json")                       # the plain ODESystem, before OEDSystem()
#   run(`python -m dynamicoed.jl_b200.ir build lv.json`)       # -> {"library": ..., "model": ...}
module DynoedIR

using ModelingToolkit
using ModelingToolkit: getdefault, hasdefault, istunable, isinput, getbounds, hasbounds
using Symbolics
using DynamicOED: is_measured, get_measurement_rate

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

function export_ir(sys::ModelingToolkit.AbstractODESystem, path::AbstractString;
        tspan = ModelingToolkit.get_tspan(sys))
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
    print(io, "]}\n")
    write(path, String(take!(io)))
    path
end

end # module
