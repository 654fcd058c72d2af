# DynareB200.jl -- the `ccall` binding a Dynare.jl maintainer adds to route the per-draw likelihood
# through libdynare_b200.so.  NOT executable in the build container (no Julia); every call below is
# mirrored 1:1 by dynare.jl_b200/_lib.py + engine.py, which is what the tests exercise.
#
# Drop-in point: `loglikelihood(θ, context, observations, ssws)` (src/estimation/estimation.jl:603-702) and
# the closure `llikelihood` built for SMC (src/estimation/estimation.jl:316-326).
module DynareB200

const LIB = get(ENV, "DYNARE_B200_LIB", "libdynare_b200.so")

struct DybError <: Exception
    code::Cint
    msg::String
end
check(rc) = rc == 0 || throw(DybError(rc, unsafe_string(ccall((:dyb_last_error, LIB), Cstring, ()))))

# estimated-parameter kinds: Dynare.EstimatedParameterType (src/dynare_containers.jl:16) -> DYB_EST_*
const KIND = Dict(:EstParameter => 0, :EstSDShock => 1, :EstVarShock => 2, :EstCorrShock => 3,
                  :EstSDMeasurement => 4, :EstVarMeasurement => 5, :EstCorrMeasurement => 6)

"Batched analogue of `SSWs` (src/estimation/estimation.jl:358-429): owns the device handle."
mutable struct BatchSSWs
    handle::Ptr{Cvoid}
    np::Int
    function BatchSSWs(context, observations::AbstractMatrix, modelname::AbstractString; device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:dyb_create, LIB), Cint, (Cstring, Cint, Ref{Ptr{Cvoid}}), modelname, device, h))
        ws = new(h[], 0)
        finalizer(w -> ccall((:dyb_destroy, LIB), Cint, (Ptr{Cvoid},), w.handle), ws)
        # calibration: context.work.params, model.Sigma_e, work.Sigma_m (src/dynare_containers.jl:846-934)
        m = context.models[1]
        check(ccall((:dyb_set_calibration, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    ws.handle, context.work.params, Matrix(m.Sigma_e), Matrix(context.work.Sigma_m)))
        # estimated-parameter map: ep.index / ep.parametertype (src/dynare_containers.jl:800-841), 1-based -> 0-based
        ep = context.work.estimated_parameters
        kind = Int32[KIND[Symbol(t)] for t in ep.parametertype]
        i1 = Int32[(x isa Pair ? x[1] : x) - 1 for x in ep.index]
        i2 = Int32[(x isa Pair ? x[2] : 1) - 1 for x in ep.index]
        check(ccall((:dyb_set_estimated_params, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                    ws.handle, length(kind), kind, i1, i2))
        ws.np = length(kind)
        # observations: ny x nobs, missing -> NaN (src/estimation/estimation.jl:603-608)
        Y = Float64[ismissing(x) ? NaN : Float64(x) for x in observations]
        check(ccall((:dyb_set_observations, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32),
                    ws.handle, Y, size(Y, 1), size(Y, 2)))
        return ws
    end
end

"""
    loglikelihood!(out, status, Θ, ws)

Batched `loglikelihood`: `Θ` is `np × N` (one draw per column), `out[j] = -Inf` wherever `status[j] != 0`,
which reproduces the try/catch of `(problem::DSGELogPosteriorDensity)(θ)` (estimation.jl:503-515).
"""
function loglikelihood!(out::Vector{Float64}, status::Vector{Int32}, Θ::Matrix{Float64}, ws::BatchSSWs)
    @assert size(Θ, 1) == ws.np && length(out) == size(Θ, 2) == length(status)
    GC.@preserve out status Θ check(ccall((:dyb_loglik_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}), ws.handle, Θ, size(Θ, 2), out, status))
    return out
end

"Scalar method, same signature role as the reference's `loglikelihood(θ, context, observations, ssws)`."
function loglikelihood(θ::AbstractVector{<:Real}, ws::BatchSSWs)
    out = Vector{Float64}(undef, 1); st = Vector{Int32}(undef, 1)
    loglikelihood!(out, st, reshape(collect(Float64, θ), :, 1), ws)
    return out[1]
end

"LRE.first_order_solver! results for N draws: g1 = [g1_1 | g1_2] (n × (k+nx) × N), ȳ (n × N)."
function first_order!(g1::Array{Float64,3}, ybar::Matrix{Float64}, status::Vector{Int32}, Θ::Matrix{Float64}, ws::BatchSSWs)
    GC.@preserve g1 ybar status Θ check(ccall((:dyb_first_order_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        ws.handle, Θ, size(Θ, 2), g1, ybar, C_NULL, C_NULL, status))
    return g1
end

# SMC primitives (src/estimation/smc.jl:118-143, 340-361); indices come back 0-based
function smc_reweight!(wout, w, loglh, dphi; device = 0)
    z = Ref(0.0); e = Ref(0.0)
    check(ccall((:dyb_smc_reweight, LIB), Cint, (Cint, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Ptr{Float64}, Ref{Float64}, Ref{Float64}),
                device, w, loglh, dphi, length(w), wout, z, e))
    return z[], e[]
end
function multinomial_resampling(w::Vector{Float64}, u::Vector{Float64}; device = 0)
    idx = Vector{Int64}(undef, length(w)); nc = Ref{Int64}(0)
    check(ccall((:dyb_smc_resample_multinomial, LIB), Cint, (Cint, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ref{Int64}),
                device, w, u, length(w), idx, nc))
    return idx .+ 1
end

end # module
