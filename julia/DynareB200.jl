# DynareB200.jl -- the `ccall` binding a Dynare.jl maintainer adds to route the per-draw likelihood
# through libdynare_b200.so.  NOT executable in the build container (no Julia); every call below is
# mirrored 1:1 by dynare.jl_b200/_lib.py + engine.py, which is what the tests exercise.
#
# Drop-in point: `loglikelihood(θ, context, observations, ssws)` (src/estimation/estimation.jl:603-702) and
# the closure `llikelihood` built for SMC (src/estimation/estimation.jl:316-326).
module DynareB200

using LinearAlgebra

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
# prox!(RPSD, IndPSD(), R); cholesky(RPSD).L (src/estimation/smc.jl:161-165); flag 0: R was positive definite, 1: projected
# first, 2: the reference's cholesky would throw
function covariance_factor(R::Matrix{Float64}; device = 0)
    np = size(R, 1); RPSD = similar(R); L = similar(R); flag = Ref{Int32}(0)
    check(ccall((:dyb_smc_covariance_factor, LIB), Cint, (Cint, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                device, R, np, RPSD, L, flag))
    flag[] == 2 && throw(PosDefException(1))
    return RPSD, LowerTriangular(L), Int(flag[])
end
function multinomial_resampling(w::Vector{Float64}, u::Vector{Float64}; device = 0)
    idx = Vector{Int64}(undef, length(w)); nc = Ref{Int64}(0)
    check(ccall((:dyb_smc_resample_multinomial, LIB), Cint, (Cint, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ref{Int64}),
                device, w, u, length(w), idx, nc))
    return idx .+ 1
end


# ---- priors, log-posterior (src/estimation/estimation.jl:455-462, 503-515) ---------------------------------------
# shape codes are the reference's own (src/estimation/priors.jl:482-529); (a, b) = params(prior) of the Distribution
const SHAPE = Dict(:Beta => 1, :Gamma => 2, :Normal => 3, :InverseGamma1 => 4, :Uniform => 5, :InverseGamma => 6, :Weibull => 8)
function set_priors!(ws::BatchSSWs, priors::Vector)
    shape = Int32[SHAPE[nameof(typeof(p))] for p in priors]
    ab = [Float64.(params(p)) for p in priors]           # Distributions.params: (a, b) / (μ, σ) / (α, θ)
    a = Float64[x[1] for x in ab]; b = Float64[x[2] for x in ab]
    check(ccall((:dyb_set_priors, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                ws.handle, length(shape), shape, a, b))
end
"`prior_draw!` (src/estimation/estimation.jl:329-335) for n draws at once on the device: np × n, column j = particle first + j."
function prior_draw(ws::BatchSSWs, n::Integer; seed = 0, first = 0, round = 0)
    para = Matrix{Float64}(undef, ws.np, n)
    check(ccall((:dyb_prior_draw, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Int32, Ptr{Float64}), ws.handle, n, first, seed, round, para))
    return para
end
"Batched `(problem::DSGELogPosteriorDensity)(θ)`: one column of Θ per draw, `-Inf` for failed draws."
function logposterior!(out::Vector{Float64}, Θ::Matrix{Float64}, ws::BatchSSWs)
    GC.@preserve out Θ check(ccall((:dyb_logposterior_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}), ws.handle, Θ, size(Θ, 2), out, C_NULL, C_NULL))
    return out
end

# ---- SMC with device-resident particles: replaces `smc(pdraw!, logprior, loglikelihood, names, np)` (smc.jl:64-313) ----
struct SmcOptions          # mirrors dyb_smc_options; defaults = Tune (smc.jl:13-62)
    npart::Int64; nphi::Int32; lam::Float64; c0::Float64; acpt0::Float64; trgt::Float64; alp::Float64
    resampling::Int32; seed::UInt64
end
function smc(ws::BatchSSWs, pdraw!::Union{Function, Nothing} = nothing; npart = 1000, nphi = 400, lam = 2.0, seed = 0)
    opt = Ref(SmcOptions(npart, nphi, lam, 0.5, 0.25, 0.25, 0.9, 0, seed))
    s = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dyb_smc_create, LIB), Cint, (Ptr{Cvoid}, Ref{SmcOptions}, Ptr{Float64}, Ref{Ptr{Cvoid}}), ws.handle, opt, C_NULL, s))
    first = Ref{Int64}(0); nloc = Ref{Int64}(0)
    check(ccall((:dyb_smc_local_range, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), s[], first, nloc))
    para = Matrix{Float64}(undef, ws.np, nloc[]); st = ones(Int32, nloc[]); θ = Vector{Float64}(undef, ws.np)
    if pdraw! === nothing                                  # prior draws + rejection loop on the device (needs set_priors!)
        check(ccall((:dyb_smc_init_from_prior, LIB), Cint, (Ptr{Cvoid}, UInt64, Int32, Ptr{Int64}), s[], seed, 1000, C_NULL))
    else
        while any(!=(0), st)                               # smc.jl:84-101: redraw until the likelihood is finite
            for j in findall(!=(0), st); pdraw!(θ); para[:, j] .= θ; end
            check(ccall((:dyb_smc_init, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}), s[], para, st))
        end
    end
    check(ccall((:dyb_smc_run, LIB), Cint, (Ptr{Cvoid}, Int32), s[], nphi))      # all stages, asynchronous
    stats = Matrix{Float64}(undef, 6, nphi); mu = Vector{Float64}(undef, ws.np); R = Matrix{Float64}(undef, ws.np, ws.np)
    logmdd = Ref(0.0)
    check(ccall((:dyb_smc_get_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}), s[], stats, mu, R, logmdd))
    w = Vector{Float64}(undef, nloc[]); loglh = similar(w); logpost = similar(w)
    check(ccall((:dyb_smc_get_particles, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s[], para, w, loglh, logpost))
    ccall((:dyb_smc_destroy, LIB), Cint, (Ptr{Cvoid},), s[])
    return (para = para, w = w, loglh = loglh, logpost = logpost, stats = stats, mu = mu, R = R, logmdd = logmdd[])
end

# ---- RWMH on many chains: replaces run_mcmc (estimation.jl:964-989); y np × nchains in the transformed space ----------
function rwmh!(y::Matrix{Float64}, ws::BatchSSWs, cholcov::Matrix{Float64}, niter::Integer; seed = 0, transformed = true)
    n = size(y, 2); logp = Vector{Float64}(undef, n); acc = zeros(Int64, n)
    GC.@preserve y check(ccall((:dyb_rwmh_run, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int32, UInt32, Int32, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        ws.handle, n, 0, niter, 0, transformed ? 1 : 0, cholcov, seed, y, logp, acc, C_NULL, C_NULL))
    return logp, acc
end

# ---- several GPUs: one Julia process per GPU (Distributed / MPI); rank 0 creates the id, everyone attaches --------------
unique_id() = (id = Vector{UInt8}(undef, 128); check(ccall((:dyb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
comm_init!(ws::BatchSSWs, rank, nranks, id::Vector{UInt8}) =
    check(ccall((:dyb_comm_init, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ws.handle, rank, nranks, id))

# ---- any model, host-evaluated Jacobians (no generated device function) ---------------------------------------------------
struct ModelDesc
    n::Int32; nx::Int32; ny::Int32; n_static::Int32; n_bkwrd_b::Int32; n_fwrd_b::Int32
    i_static::Ptr{Int32}; i_bkwrd_b::Ptr{Int32}; i_fwrd_b::Ptr{Int32}; obs_idx::Ptr{Int32}
end
"jac: n × ncol × N from `set_compressed_jacobian!` per draw (threaded on the host), ybar: n × N."
function loglik_from_jacobian!(out, status, h::Ptr{Cvoid}, jac::Array{Float64,3}, ybar::Matrix{Float64}, Σe::Matrix{Float64})
    GC.@preserve out status jac ybar Σe check(ccall((:dyb_loglik_from_jacobian_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Int32}),
        h, jac, ybar, Σe, 0, C_NULL, 0, C_NULL, size(jac, 3), out, status))
    return out
end

"Handle for a model without a generated device function (index sets as `Model` holds them, 1-based here)."
function create_generic(n, nx, i_static, i_bkwrd_b, i_fwrd_b, obs_idx; device = 0)
    a = Int32.(i_static .- 1); b = Int32.(i_bkwrd_b .- 1); c = Int32.(i_fwrd_b .- 1); o = Int32.(obs_idx .- 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve a b c o begin
        desc = Ref(ModelDesc(n, nx, length(o), length(a), length(b), length(c), pointer(a), pointer(b), pointer(c), pointer(o)))
        check(ccall((:dyb_create_generic, LIB), Cint, (Ref{ModelDesc}, Cint, Ref{Ptr{Cvoid}}), desc, device, h))
    end
    return h[]
end

# ---- prior-predictive quantities and the smoother for a batch of draws ---------------------------------------------
"run_simulations (priorprediction.jl:64-98) for all columns of Θ: steady state, stdev, IRFs (n × periods × nx × N)."
function prior_predictive(ws::BatchSSWs, Θ::Matrix{Float64}, n, nx; periods = 40)
    N = size(Θ, 2)
    ybar = Matrix{Float64}(undef, n, N); sd = similar(ybar); irfs = Array{Float64,4}(undef, n, periods, nx, N)
    st = Vector{Int32}(undef, N)
    GC.@preserve Θ check(ccall((:dyb_prior_predictive_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        ws.handle, Θ, N, periods, ybar, sd, irfs, st))
    return ybar, sd, irfs, st
end
"calibsmoother! (kalman.jl:52-264) for all columns of Θ: smoothed variables in levels (n × nobs × N), smoothed shocks."
function smoother(ws::BatchSSWs, Θ::Matrix{Float64}, n, nx, ny, nobs)
    N = size(Θ, 2)
    alphah = Array{Float64,3}(undef, n, nobs, N); etah = Array{Float64,3}(undef, nx, nobs, N)
    epsh = Array{Float64,3}(undef, ny, nobs, N); ll = Vector{Float64}(undef, N); st = Vector{Int32}(undef, N)
    GC.@preserve Θ check(ccall((:dyb_smoother_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        ws.handle, Θ, N, alphah, etah, epsh, ll, st))
    return alphah, etah, epsh, ll, st
end
"Σ_e (nx × nx × N) and Σ_m (ny × ny × N) of every column of Θ, as set_estimated_parameters! builds them (estimation.jl:517-601)."
function shock_covariances(ws::BatchSSWs, Θ::Matrix{Float64}, nx, ny)
    N = size(Θ, 2)
    se = Array{Float64,3}(undef, nx, nx, N); sm = Array{Float64,3}(undef, ny, ny, N)
    GC.@preserve Θ check(ccall((:dyb_shock_covariances_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}), ws.handle, Θ, N, se, sm))
    return se, sm
end
"""
The non-stationary branch of calibsmoother! (kalman.jl:163-223) for N state spaces sharing one data set.  `T` ns×ns×N,
`R` ns×nx×N, `Q` nx×nx×N; `Pinf0`, `Pstar0` ns×ns×N in the ORIGINAL basis: with `F = Schur(gees!(...))` as at
kalman.jl:166-174 and `k` unit roots, `Pinf0 = F.Z[:, 1:k] * F.Z[:, 1:k]'`, `Pstar0 = F.Z[:, k+1:end] * P22 * F.Z[:, k+1:end]'`.
`z_idx` 1-based observed states, `Y` ny×nobs with NaN for missing, `Hdiag` ny×N or nothing.
"""
function diffuse_smoother(Y::Matrix{Float64}, z_idx::Vector{Int}, T::Array{Float64,3}, R::Array{Float64,3},
                          Q::Array{Float64,3}, Pinf0::Array{Float64,3}, Pstar0::Array{Float64,3};
                          Hdiag = nothing, ybar_obs = nothing, tol = 1e-8, device = 0)
    ns, _, N = size(T); nx = size(R, 2); ny, nobs = size(Y)
    z0 = Int32.(z_idx .- 1)
    alphah = Array{Float64,3}(undef, ns, nobs, N); etah = Array{Float64,3}(undef, nx, nobs, N)
    epsh = Array{Float64,3}(undef, ny, nobs, N); att = similar(alphah); apred = Array{Float64,3}(undef, ns, nobs + 1, N)
    ll = Vector{Float64}(undef, N); nd = Vector{Int32}(undef, N); st = Vector{Int32}(undef, N)
    hd = Hdiag === nothing ? Ptr{Float64}(C_NULL) : pointer(Hdiag)
    yb = ybar_obs === nothing ? Ptr{Float64}(C_NULL) : pointer(ybar_obs)      # ny × N, subtracted from Y
    GC.@preserve Y z0 T R Q Pinf0 Pstar0 Hdiag ybar_obs check(ccall((:dyb_kalman_diffuse_smoother_batch, LIB), Cint,
        (Cint, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
        device, ns, ny, nx, nobs, z0, Y, T, R, Q, Pinf0, Pstar0, C_NULL, hd, yb, tol, N, alphah, etah, epsh, att, apred, ll, nd, st))
    return (; alphah, etah, epsilonh = epsh, att, a_pred = apred, loglik = ll, n_diffuse = nd, status = st)
end
"FiniteDiff.finite_difference_hessian(objective, x) (estimation.jl:873, 882) as ONE batched posterior evaluation."
function finite_difference_hessian(ws::BatchSSWs, x::Vector{Float64})
    n = length(x); h = max.(eps()^(1/4) .* abs.(x), eps()^(1/4))
    pts = [copy(x)]
    for i in 1:n, m in (2.0, 1.0, -1.0, -2.0); p = copy(x); p[i] += m * h[i]; push!(pts, p); end
    for i in 1:n, j in i+1:n, (si, sj) in ((1, 1), (1, -1), (-1, 1), (-1, -1)); p = copy(x); p[i] += si * h[i]; p[j] += sj * h[j]; push!(pts, p); end
    f = -logposterior!(Vector{Float64}(undef, length(pts)), reduce(hcat, pts), ws)
    H = Matrix{Float64}(undef, n, n); k = 2
    for i in 1:n
        H[i, i] = (-f[k] + 16f[k+1] - 30f[1] + 16f[k+2] - f[k+3]) / (12h[i]^2); k += 4
    end
    for i in 1:n, j in i+1:n
        H[i, j] = H[j, i] = (f[k] - f[k+1] - f[k+2] + f[k+3]) / (4h[i] * h[j]); k += 4
    end
    return H
end


# ---- the estimation ENTRY POINTS on `context` (mirrored 1:1 by dynare_jl_b200/estimation.py) ------------------------------
# A Dynare.jl maintainer adds these as methods / replaces the bodies of the functions of the same names.
#
# device_model!(context): make the model of `context` known to the library.  The preprocessor has already written
# <modfilepath>/model/json/modfile.json and <modfilepath>/model/julia/*.jl (that is what `context` was built from,
# src/DynareParser.jl:10-15, 169-218; src/dynare_functions.jl:10-47); dynare_jl_b200.preprocessor translates exactly those
# files into the CUDA model function, nvcc compiles it (~15 s, cached next to the model) and the shared object registers
# the model under `name` when it is dlopen'ed.
const MODEL_LIBS = Dict{String,Ptr{Cvoid}}()
function device_model!(context; name = basename(context.modfileinfo.modfilepath),
                       python = get(ENV, "DYNARE_B200_PYTHON", "python3"))
    haskey(MODEL_LIBS, name) && return name
    Libdl = Base.require(Base.PkgId(Base.UUID("8f399da3-3557-5675-b5ff-fb832c97cbdb"), "Libdl"))
    Libdl.dlopen(LIB, Libdl.RTLD_GLOBAL | Libdl.RTLD_NOW)
    outdir = joinpath(context.modfileinfo.modfilepath, "model", "b200")
    lib = joinpath(outdir, "libdynare_b200_model_$(name).so")
    if !isfile(lib) || mtime(lib) < mtime(joinpath(context.modfileinfo.modfilepath, "model", "json", "modfile.json"))
        run(`$python -m dynare_jl_b200.preprocessor $name $(context.modfileinfo.modfilepath) $outdir`)
    end
    MODEL_LIBS[name] = Libdl.dlopen(lib, Libdl.RTLD_GLOBAL | Libdl.RTLD_NOW)
    return name
end

"`SSWs(context, observations, varobs)` replacement: device model + calibration + estimated-parameter map + priors."
function BatchSSWs(context, observations::AbstractMatrix; device::Integer = 0)
    ws = BatchSSWs(context, observations, device_model!(context); device = device)
    set_priors!(ws, context.work.estimated_parameters.prior)
    return ws
end

"""
    smc_compute!(; context, datafile = "", first_obs = 1, last_obs = 0, nobs = 0, npart = 1000, nphi = 400, lam = 2.0, seed = 0)

Replaces the body of `smc_compute!` (src/estimation/estimation.jl:298-327): the three closures it builds for `smc`
(`pdraw!`, `logprior`, `llikelihood`) disappear -- prior density, likelihood, reweighting, resampling and mutation all run
on the device; only `prior_draw!(θ, ep)` (estimation.jl:329-333) stays on the host for the initial particles.
"""
function smc_compute!(; context, datafile = "", data = nothing, first_obs = 1, last_obs = 0, nobs = 0,
                      npart = 1000, nphi = 400, lam = 2.0, seed = 0, device = 0)
    ep = context.work.estimated_parameters
    observations = Main.Dynare.get_observations(context, datafile, data, first_obs, last_obs, nobs)
    ws = BatchSSWs(context, observations; device = device)
    pdraw!(θ) = Main.Dynare.prior_draw!(θ, ep)
    r = smc(ws, pdraw!; npart = npart, nphi = nphi, lam = lam, seed = seed)
    μ = r.para * r.w                                        # posterior mean (smc.jl:283-290)
    σ = sqrt.(max.(((r.para .- μ) .^ 2) * r.w, 0.0))
    return (; r..., posterior_mean = μ, posterior_std = σ)
end

"""
    rwmh_compute!(; context, datafile, data, initial_values, covariance, transformed_covariance, mcmc_chains, mcmc_jscale,
                  mcmc_replic, transformed_parameters = true, seed = 0, ...)

Replaces `rwmh_compute!` -> `mh_estimation` -> `run_mcmc` (estimation.jl:233-279, 912-989): all `mcmc_chains` chains advance
together, one batched posterior evaluation per sweep (the reference runs them on `MCMCThreads()` over ONE shared mutable
`context`).  Returns the chain histories in parameter space (np × mcmc_chains × mcmc_replic) and the acceptance rates.
"""
function rwmh_compute!(; context, datafile = "", data = nothing, first_obs = 1, last_obs = 0, nobs = 0,
                       initial_values = Main.Dynare.prior_mean(context.work.estimated_parameters),
                       covariance = Matrix(Main.Dynare.prior_variance(context.work.estimated_parameters)),
                       transformed_covariance = Matrix{Float64}(undef, 0, 0), mcmc_chains = 1, mcmc_jscale = 0.0,
                       mcmc_replic = 0, transformed_parameters = true, seed = 0, device = 0)
    observations = Main.Dynare.get_transposed_data!(context, datafile, data, context.work.observed_variables, first_obs, last_obs, nobs)
    ws = BatchSSWs(context, observations; device = device)
    np = ws.np
    y0 = reshape(collect(Float64, initial_values), np, 1)
    if transformed_parameters
        y0 = transform(ws, y0; inverse = true)
        Σ = isempty(transformed_covariance) ? covariance : transformed_covariance
    else
        Σ = covariance
    end
    Σ = mcmc_jscale .* (Σ .+ Σ') ./ 2                       # run_mcmc symmetrises (estimation.jl:966)
    L = Matrix(cholesky(Σ).L)
    y = repeat(y0, 1, mcmc_chains)
    hist = Array{Float64,3}(undef, np, mcmc_chains, mcmc_replic); hlp = Matrix{Float64}(undef, mcmc_chains, mcmc_replic)
    logp = Vector{Float64}(undef, mcmc_chains); acc = zeros(Int64, mcmc_chains)
    GC.@preserve y L hist hlp check(ccall((:dyb_rwmh_run, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int32, UInt32, Int32, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        ws.handle, mcmc_chains, 0, mcmc_replic, 0, transformed_parameters ? 1 : 0, L, seed, y, logp, acc, hist, hlp))
    if transformed_parameters                              # transform_chains! (estimation.jl:957-959)
        hist = reshape(transform(ws, reshape(hist, np, :)), np, mcmc_chains, mcmc_replic)
    end
    return (chains = hist, logp = hlp, acceptance_rate = acc ./ max(mcmc_replic, 1))
end

"DSGETransformation (estimation.jl:465-487) for all columns of X: ℝ^np -> parameters, or back with `inverse = true`."
function transform(ws::BatchSSWs, X::Matrix{Float64}; inverse = false)
    out = similar(X); lj = Vector{Float64}(undef, size(X, 2))
    GC.@preserve X check(ccall((:dyb_transform_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}),
                                ws.handle, X, size(X, 2), inverse ? 1 : 0, out, lj))
    return out
end

"""
    posterior_mode_hessians!(context, ws, minimizer)

The two `finite_difference_hessian` calls of `posterior_mode!` (estimation.jl:873-890) -- at the transformed minimiser and at
the mode in parameter space -- each as ONE batch of 1 + 4 np + 2 np (np - 1) posterior evaluations; fills the same result
fields.  The optimiser itself (`dynare_optimize`, estimation.jl:864-872) keeps calling the scalar `loglikelihood(θ, ws)`.
"""
function posterior_mode_hessians!(context, ws::BatchSSWs, mode_parameters::Vector{Float64})
    results = context.results.model_results[1].estimation
    hess = finite_difference_hessian(ws, mode_parameters)
    hsd = sqrt.(diag(hess))
    invhess = inv(hess ./ (hsd * hsd')) ./ (hsd * hsd')
    d = diag(invhess); d[d .< 0] .= NaN
    results.posterior_mode = copy(mode_parameters)
    results.posterior_mode_std = sqrt.(d)
    results.posterior_mode_covariance = invhess
    return results
end

"`run_simulations(context, draws)` (priorprediction.jl:64-98) for all rows of `draws` in one call; failure classes from the statuses."
function run_simulations(context, draws::Matrix{Float64}; device = 0)
    m = context.models[1]
    ws = BatchSSWs(context, Matrix{Float64}(undef, length(context.work.observed_variables), 0); device = device)
    ybar, sd, irfs, st = prior_predictive(ws, Matrix(draws'), m.endogenous_nbr, m.exogenous_nbr)
    failure = Float64.(st .!= 0)
    return failure, (steady_state = ybar, stdev = sd, irfs = irfs, domain = count(==(1), st), undetermined = count(==(2), st),
                     unstable = count(==(3), st), other = count(>(3), st))
end

end # module
