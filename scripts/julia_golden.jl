# Reference-pinned golden vectors: evaluates the REAL Dynare.jl hot path on the inputs written by
# scripts/make_ref_input.py and stores what it returns, so that tests/test_reference_pinned.py can compare the CPU oracle
# and the CUDA engine with the reference itself instead of with our reading of it.
#
#   julia --project=<env with Dynare.jl v0.10.x> scripts/julia_golden.jl <path to the Dynare.jl checkout> [model ...]
#
# For every model (default: fs2000 ls2003 hs_nk) and every parameter vector theta of tests/golden/ref_input_<model>.json:
#   * Dynare.loglikelihood(theta, context, observations, ssws)          (src/estimation/estimation.jl:603-702)
#     inside the same try/catch as DSGELogPosteriorDensity              (src/estimation/estimation.jl:503-515);
#     the exception type is recorded so that the engine's per-draw status codes can be checked too
#   * the first-order solution LRE.first_order_solver! left in context.results.model_results[1]
#     .linearrationalexpectations (g1_1, g1_2; src/perturbations.jl:663-664) and the steady state
#   * logpriordensity(theta, estimated_parameters)                      (src/estimation/estimation.jl:455-462)
# Output: tests/golden/ref_<model>.json (committed; a few hundred KB).
#
# The .mod files are the reference's own (test/models/...); the estimation / stoch_simul commands at their end are
# dropped from a scratch copy because they need data files that are not in the tree -- the observations come from the
# input JSON instead (for ls2003 they ARE the reference's data_ca1.csv rows 8..86).
using Dynare
using JSON
using LinearAlgebra

const REPO = normpath(joinpath(@__DIR__, ".."))
const GOLD = joinpath(REPO, "tests", "golden")
const DROP = ("estimation(", "stoch_simul(", "calib_smoother(", "mode_compute!(", "rwmh_compute!(", "smc_compute!(",
              "identification", "check;", "forecast(")

function scratch_modfile(src::String, workdir::String)
    dst = joinpath(workdir, basename(src))
    open(dst, "w") do io
        for ln in eachline(src)
            s = lstrip(ln)
            any(startswith(s, d) for d in DROP) && continue
            println(io, ln)
        end
    end
    # data files referenced by name are not needed any more, but auxiliary files beside the model may be
    for f in readdir(dirname(src))
        p = joinpath(dirname(src), f)
        (isfile(p) && !endswith(f, ".mod")) && cp(p, joinpath(workdir, f); force = true)
    end
    return dst
end

tojson(x::AbstractMatrix) = [collect(Float64, x[i, :]) for i in axes(x, 1)]      # row lists

function run_model(dynare_root::String, name::String)
    inp = JSON.parsefile(joinpath(GOLD, "ref_input_$(name).json"))
    workdir = mktempdir()
    modfile = scratch_modfile(joinpath(dynare_root, inp["modfile"]), workdir)
    context = Dynare.dynare(modfile)
    varobs = context.work.observed_variables
    @assert String.(varobs) == String.(inp["varobs"]) "varobs order differs: $(varobs) vs $(inp["varobs"])"
    ny = length(inp["Y"]); nobs = length(inp["Y"][1])
    observations = Matrix{Union{Missing,Float64}}(undef, ny, nobs)
    for i in 1:ny, t in 1:nobs
        v = inp["Y"][i][t]
        observations[i, t] = v === nothing ? missing : Float64(v)
    end
    ssws = Dynare.SSWs(context, observations, varobs)
    ep = context.work.estimated_parameters
    results = context.results.model_results[1]
    lre = results.linearrationalexpectations
    out = Dict{String,Any}(
        "model" => name, "dynare_version" => string(pkgversion(Dynare)), "julia_version" => string(VERSION),
        "estimated" => String.(string.(ep.name)), "state_ids" => collect(ssws.state_ids),
        "i_bkwrd_b" => collect(context.models[1].i_bkwrd_b),
        "loglik" => Any[], "logprior" => Any[], "error" => Any[], "g1_1" => Any[], "g1_2" => Any[], "steady_state" => Any[],
    )
    for th in inp["theta"]
        θ = Float64.(th)
        ll = -Inf; err = ""
        try
            ll = Dynare.loglikelihood(θ, context, observations, ssws)
        catch e
            err = string(typeof(e))
            ll = -Inf
        end
        lp = try Dynare.logpriordensity(θ, ep) catch; -Inf end
        push!(out["loglik"], isfinite(ll) ? ll : nothing)
        push!(out["logprior"], isfinite(lp) ? lp : nothing)
        push!(out["error"], err)
        if err == ""
            push!(out["g1_1"], tojson(lre.g1_1)); push!(out["g1_2"], tojson(lre.g1_2))
            push!(out["steady_state"], collect(Float64, results.trends.endogenous_steady_state))
        else
            push!(out["g1_1"], nothing); push!(out["g1_2"], nothing); push!(out["steady_state"], nothing)
        end
    end
    open(joinpath(GOLD, "ref_$(name).json"), "w") do io
        JSON.print(io, out)
    end
    nok = count(==(""), out["error"])
    println("$(name): $(length(inp["theta"])) draws, $(nok) evaluated, written tests/golden/ref_$(name).json")
end

function main()
    isempty(ARGS) && error("usage: julia scripts/julia_golden.jl <Dynare.jl checkout> [model ...]")
    models = length(ARGS) > 1 ? ARGS[2:end] : ["fs2000", "ls2003", "hs_nk"]
    for m in models
        run_model(ARGS[1], m)
    end
end

main()
