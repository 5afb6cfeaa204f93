# DataDrivenB200.jl — Julia glue for libddb200.so (C ABI in include/ddb.h).
#
# UNTESTED IN THIS REPOSITORY: neither the build container nor the GPU box has a Julia toolchain
# (probed: no `julia` binary, no ~/.julia).  The file mirrors, call for call, the Python host layer
# `datadrivendiffeq.jl_b200/api.py` + `backend.py`, which IS tested against the same shared library.
#
# Usage (drop-in, the reference's `solve` API is unchanged):
#
#     using DataDrivenDiffEq, DataDrivenSparse, DataDrivenB200
#     res = solve(prob, basis, B200Sparse(STLSQ(exp10.(-5:0.1:-1))))                # one GPU
#     res = solve(prob, basis, B200Sparse(STLSQ(exp10.(-5:0.1:-1)); ngpus = 8))     # the node's eight GPUs
#
# `ngpus > 1`: ONE Julia process drives N GPUs through a `ddb_group` (N contexts + `ncclCommInitAll` inside the
# library); `ddb_group_solve_sparse` cuts X / Y into N sample ranges, each GPU contracts its own range, the packed
# normal-equation blocks and the candidate-rss table are summed with one `ncclAllReduce` each on the compute streams.
#
# The backend is a new algorithm type.  It defines exactly the two methods the reference's plug-in
# mechanism dispatches on (SURVEY.md section 8b): `get_fit_targets` (src/commonsolve.jl:74-81) so that
# `init` never evaluates the basis on the data, and `solve!` (more specific than
# lib/DataDrivenSparse/src/commonsolve.jl:1-5).
module DataDrivenB200

using DataDrivenDiffEq, DataDrivenSparse, CommonSolve
using DataDrivenDiffEq: InternalDataDrivenProblem, DataDrivenSolution, DDReturnCode, get_implicit_data,
                        get_oop_args, __construct_basis, states, equations
using DataDrivenDiffEq: DataDrivenCommonOptions, DataNormalization, DataProcessing
import Symbolics
import LinearAlgebra
import Random
import Parameters
import StatsBase: ZScoreTransform, UnitRangeTransform

const LIB = get(ENV, "DDB200_LIB", joinpath(@__DIR__, "..", "libddb200.so"))

struct SweepParams            # ddb_sweep_params
    algorithm::Int32; proximal::Int32; selector::Int32; maxiters::Int32
    abstol::Float64; reltol::Float64; rho::Float64; cad_rho::Float64
end

struct B200Sparse{A <: Union{STLSQ, ADMM, SR3, ImplicitOptimizer}} <: DataDrivenSparse.AbstractSparseRegressionAlgorithm
    inner::A
    device::Int
    ngpus::Int
end
B200Sparse(inner; device::Int = 0, ngpus::Int = 1) = B200Sparse(inner, device, ngpus)
B200Sparse(inner, device::Int) = B200Sparse(inner, device, 1)
Base.summary(x::B200Sparse) = "B200(" * summary(x.inner) * ")"

function check(status::Cint)
    status == 0 && return
    msg = unsafe_string(ccall((:ddb_last_error, LIB), Cstring, ()))
    throw(ErrorException("ddb status $status: $msg"))   # never a silent CPU fallback
end

# HOOK 1: hand `init` the RAW states and targets; Theta is never built (the default
# DataNormalization{Nothing} then is a no-op, src/utils/data_processing.jl:64-66,114-115).
function DataDrivenDiffEq.get_fit_targets(::B200Sparse, prob::DataDrivenDiffEq.AbstractDataDrivenProblem,
        basis::DataDrivenDiffEq.AbstractBasis)
    X = first(get_oop_args(prob))
    Y = get_implicit_data(prob)
    return X, Y
end

# HOOK 0: `init` (src/commonsolve.jl:92-139).  The reference's `init` fits the normalisation on what get_fit_targets
# returns, APPLIES it in place and cuts the batches out of it -- all of which is meant for Theta, not for the raw states
# this backend hands over.  So the generic `init` runs with neutral pre-processing options (identity transform, one
# batch: the data stay where they are, nothing is copied or scaled) and the internal problem then carries the CALLER'S
# options; normalisation, train / test split and shuffled mini-batches are done on the resident data by `solve!` below
# (`solve_processed`), call for call like `api.py: _solve_processed`.
function CommonSolve.init(prob::DataDrivenDiffEq.AbstractDataDrivenProblem, basis::DataDrivenDiffEq.AbstractBasis,
        alg::B200Sparse; options::DataDrivenCommonOptions = DataDrivenCommonOptions(), kwargs...)
    options.denoise && throw(ArgumentError("B200Sparse: denoise = true needs Theta materialised and is not supported"))
    neutral = Parameters.reconstruct(options; normalize = DataNormalization(), data_processing = DataProcessing())
    ps = invoke(CommonSolve.init,
        Tuple{DataDrivenDiffEq.AbstractDataDrivenProblem, DataDrivenDiffEq.AbstractBasis,
              DataDrivenDiffEq.AbstractDataDrivenAlgorithm},
        prob, basis, alg; options = neutral, kwargs...)
    return InternalDataDrivenProblem(ps.alg, ps.testdata, ps.traindata, ps.transform, ps.control_idx, ps.implicit_idx,
        ps.parameter_idx, ps.state_idx, options, ps.basis, ps.problem, ps.kwargs)
end

needs_processing(o::DataDrivenCommonOptions) = !(o.normalize isa DataNormalization{Nothing}) ||
    o.data_processing.split < 1.0 || o.data_processing.batchsize > 0

# exponent table of a monomial basis, n x k, in the basis' own term order
function exponent_table(basis; implicits::Bool = false)
    xs = implicits ? vcat(states(basis), DataDrivenDiffEq.implicit_variables(basis)) : states(basis)
    eqs = [eq.rhs for eq in equations(basis)]
    E = zeros(UInt8, length(xs), length(eqs))
    for (j, term) in enumerate(eqs), (i, x) in enumerate(xs)
        d = Symbolics.degree(term, x)
        E[i, j] = d
    end
    # reject anything that is not a pure monomial of the states
    for (j, term) in enumerate(eqs)
        rebuilt = prod(xs[i]^Int(E[i, j]) for i in eachindex(xs); init = Symbolics.Num(1))
        isequal(Symbolics.simplify(term - rebuilt), 0) ||
            throw(ArgumentError("B200Sparse supports monomial library terms only; term $j is $(term)"))
    end
    return E
end

prox_id(::SoftThreshold) = Int32(0)
prox_id(::HardThreshold) = Int32(1)
prox_id(::ClippedAbsoluteDeviation) = Int32(2)
alg_id(::STLSQ) = Int32(0); alg_id(::ADMM) = Int32(1); alg_id(::SR3) = Int32(2)
relax(a::STLSQ) = Float64(a.rho); relax(a::ADMM) = Float64(a.rho); relax(a::SR3) = Float64(a.nu)
selector_id(f) = f === DataDrivenDiffEq.bic ? Int32(0) : f === DataDrivenDiffEq.aic ? Int32(1) :
                 f === DataDrivenDiffEq.aicc ? Int32(2) : f === DataDrivenDiffEq.rss ? Int32(3) :
                 throw(ArgumentError("B200Sparse: selector must be bic, aic, aicc or rss"))

# StatsAPI aicc of a sparse-regression cache (lib/DataDrivenSparse/src/DataDrivenSparse.jl:171-201)
function aicc_of(rss, dof, nobs)
    ll = -nobs / 2 * log(rss / nobs)
    return -2ll + 2dof + 2dof * (dof + 1) / (nobs - dof - 1)
end

# HOOK 2b: ImplicitOptimizer (lib/DataDrivenSparse/src/commonsolve.jl:58-114 + algorithms/Implicit.jl:57-108).
# The library is evaluated on (implicits = DX, states = X): the exponent table covers [states; implicits] and
# the device "states" are vcat(X, DX).  One Gram pass; then, per implicit variable, ONE batched sweep regresses
# every candidate row on the others (ddb_implicit_select) and the best left-hand side is chosen here exactly
# as the reference does (dof >= 2, touches a necessary term, smallest AICc).
function CommonSolve.solve!(ps::InternalDataDrivenProblem{<:B200Sparse{<:ImplicitOptimizer}})
    (; alg, basis, problem, options, implicit_idx) = ps
    @assert DataDrivenDiffEq.is_implicit(basis) "The provided `Basis` does not have implicit variables!"
    needs_processing(options) &&
        throw(ArgumentError("B200Sparse: normalisation / split / batches are not supported with an ImplicitOptimizer"))
    X, Y = DataDrivenDiffEq.get_fit_targets(alg, problem, basis)   # the raw data themselves (no shuffled copy)
    XA = Matrix{Float64}(vcat(X, Y)); Y = convert(Matrix{Float64}, Y)
    n, m = size(XA); n_t = size(Y, 1)
    E = exponent_table(basis; implicits = true); k = size(E, 2)
    inner = alg.inner.optimizer
    lambdas = collect(Float64, DataDrivenSparse.get_thresholds(inner))
    prox = DataDrivenSparse.get_proximal(inner)
    prm = Ref(SweepParams(alg_id(inner), prox_id(prox), selector_id(options.selector), options.maxiters,
        options.abstol, options.reltol, relax(inner), prox isa ClippedAbsoluteDeviation ? Float64(prox.ρ) : NaN))
    C = zeros(size(implicit_idx, 2), k); lams = Float64[]; its = Int[]; trainerror = 0.0
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddb_ctx_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), alg.device, ctx))
    try
        GC.@preserve XA Y E lambdas begin
            check(ccall((:ddb_set_library_monomial, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx[], n, k, E))
            check(ccall((:ddb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Int64), ctx[], XA, Y, n_t, m))
            check(ccall((:ddb_gram, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], C_NULL))
            for i in axes(implicit_idx, 2)
                others = [j for j in axes(implicit_idx, 2) if j != i]
                idx = [r for r in axes(implicit_idx, 1) if implicit_idx[r, i] || !any(implicit_idx[r, others])]
                kk = length(idx); idx0 = Int32.(idx .- 1)
                coef = zeros(kk, kk); lam = zeros(kk); iters = zeros(Int32, kk); rss = zeros(kk); dof = zeros(Int32, kk)
                GC.@preserve idx0 coef lam iters rss dof begin
                    check(ccall((:ddb_implicit_select, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint), ctx[], idx0, kk))
                    check(ccall((:ddb_sweep, LIB), Cint, (Ptr{Cvoid}, Ptr{SweepParams}, Ptr{Float64}, Cint, Int64),
                        ctx[], prm, lambdas, length(lambdas), m))
                    check(ccall((:ddb_rss, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], C_NULL))
                    check(ccall((:ddb_select, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{UInt32},
                         Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                        ctx[], coef, lam, iters, rss, dof, C_NULL, C_NULL, C_NULL, C_NULL))
                end
                necessary = implicit_idx[idx, i]
                best = 0
                for e in 1:kk
                    inds = setdiff(1:kk, e)
                    if dof[e] >= 2 && any(!iszero, coef[e, inds[necessary[inds]]])
                        if best == 0 || aicc_of(rss[e], dof[e], m) < aicc_of(rss[best], dof[best], m)
                            best = e
                        end
                    end
                end
                row = coef[best, :]; row[best] = -1.0          # BoundsError for best == 0, as in the reference
                C[i, idx] .= row
                push!(lams, lam[best]); push!(its, iters[best]); trainerror += rss[best]
            end
        end
    finally
        ccall((:ddb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), ctx[])
    end
    result = DataDrivenSparse.SparseRegressionResult(C, sum(abs.(C) .> 0.0), lams, its, nothing, trainerror, DDReturnCode(1))
    new_basis = __construct_basis(copy(C), basis, problem, options)
    return DataDrivenSolution(new_basis, problem, alg, [result], ps, result.retcode)
end

# HOOK 2
function CommonSolve.solve!(ps::InternalDataDrivenProblem{<:B200Sparse})
    (; alg, basis, problem, options) = ps
    options.denoise && throw(ArgumentError("B200Sparse: denoise=true is not supported"))
    # the raw data themselves: with the neutral `init` above `traindata` is one batch of ALL samples, and taking it from
    # the DataLoader would only add a shuffled host copy (the column order is irrelevant to the Gram sums)
    X, Y = DataDrivenDiffEq.get_fit_targets(alg, problem, basis)
    X = convert(Matrix{Float64}, X); Y = convert(Matrix{Float64}, Y)   # no copy when they already are
    n, m = size(X); n_t = size(Y, 1)
    E = exponent_table(basis); k = size(E, 2)
    inner = alg.inner
    lambdas = collect(Float64, DataDrivenSparse.get_thresholds(inner))
    prox = DataDrivenSparse.get_proximal(inner)
    prm = Ref(SweepParams(alg_id(inner), prox_id(prox), selector_id(options.selector), options.maxiters,
        options.abstol, options.reltol, relax(inner), prox isa ClippedAbsoluteDeviation ? Float64(prox.ρ) : NaN))
    if needs_processing(options)
        alg.ngpus > 1 && throw(ArgumentError("B200Sparse: normalisation / split / batches with ngpus > 1 run through the " *
            "per-rank entry points (`ddb_comm_*`, api.py: _solve_processed_sharded), not through a ddb_group"))
        return solve_processed(ps, X, Y, E, prm, lambdas)
    end
    coef = zeros(n_t, k); lam = zeros(n_t); iters = zeros(Int32, n_t); rss = zeros(n_t); dof = zeros(Int32, n_t)
    flags = zeros(Int32, n_t)
    if alg.ngpus > 1
        # one process, N GPUs: contexts + NCCL communicators live in the group (ddb.h, "multi-GPU")
        grp = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ddb_group_create, LIB), Cint, (Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}), alg.ngpus, C_NULL, grp))
        try
            GC.@preserve X Y E lambdas coef lam iters rss dof flags begin
                check(ccall((:ddb_group_set_library_monomial, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), grp[], n, k, E))
                check(ccall((:ddb_group_solve_sparse, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{SweepParams}, Ptr{Float64}, Cint,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{UInt32}, Ptr{Int32}),
                    grp[], X, Y, n_t, m, prm, lambdas, length(lambdas), coef, lam, iters, rss, dof, C_NULL, flags))
            end
        finally
            ccall((:ddb_group_destroy, LIB), Cint, (Ptr{Cvoid},), grp[])
        end
    else
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ddb_ctx_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), alg.device, ctx))
        try
            GC.@preserve X Y E lambdas coef lam iters rss dof flags begin
                check(ccall((:ddb_set_library_monomial, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx[], n, k, E))
                check(ccall((:ddb_solve_sparse, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{SweepParams}, Ptr{Float64}, Cint,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{UInt32}, Ptr{Int32}),
                    ctx[], X, Y, n_t, m, prm, lambdas, length(lambdas), coef, lam, iters, rss, dof, C_NULL, flags))
            end
        finally
            ccall((:ddb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), ctx[])
        end
    end
    # finish exactly like the reference (lib/DataDrivenSparse/src/commonsolve.jl:13-24, 37-55); a target whose
    # factorisation broke down (flags bit 0) maps to DDReturnCode(2) (src/DataDrivenDiffEq.jl:60-67)
    retcode = any(!iszero, flags .& Int32(1)) ? DDReturnCode(2) : DDReturnCode(1)
    result = DataDrivenSparse.SparseRegressionResult(coef, sum(abs.(coef) .> 0.0), lam, Int.(iters), nothing,
        sum(rss), retcode)
    new_basis = __construct_basis(copy(coef), basis, problem, options)
    return DataDrivenSolution(new_basis, problem, alg, [result], ps, result.retcode)
end

# `solve!` with the reference's pre-processing options on the RESIDENT data (src/commonsolve.jl:128-139,
# src/utils/data_processing.jl:29-39, 64-80, lib/DataDrivenSparse/src/commonsolve.jl:1-55): the transform is fitted on the
# library rows of ALL samples without building them (`ddb_term_css` / `ddb_term_minmax`), the contiguous train / test
# split and the shuffled mini-batches are views (`ddb_set_sample_range`) of the data after one device-side gather
# (`ddb_permute_samples`), every batch is regressed in the transformed space (`ddb_gram_scale_terms` /
# `ddb_gram_affine_terms` act on the normal-equation blocks), test errors come from `ddb_residual(_affine)`, and the
# result with the smallest test (or train) error wins.  Mirrors `api.py: _solve_processed` call for call.
function solve_processed(ps, X::Matrix{Float64}, Y::Matrix{Float64}, E::Matrix{UInt8}, prm, lambdas::Vector{Float64})
    (; alg, basis, problem, options) = ps
    dp = options.data_processing
    n, m = size(X); n_t = size(Y, 1); k = size(E, 2); L = length(lambdas)
    split = clamp(Float64(dp.split), 0.0, 1.0)
    m_train = clamp(round(Int, split * m), 0, m)                 # MLUtils.splitobs(at = split, shuffle = false)
    m_train > 0 || throw(ArgumentError("B200Sparse: the training split is empty"))
    unit = options.normalize isa DataNormalization{UnitRangeTransform}
    zs = options.normalize isa DataNormalization{ZScoreTransform}
    # UnitRange: a row of ones rides along as one more target, R[:, end] = Theta * 1 of every batch view
    Yd = unit ? vcat(Y, ones(1, m)) : Y
    ntd = size(Yd, 1)
    sigma = ones(k); umin = zeros(k); udiv = ones(k)
    results = DataDrivenSparse.SparseRegressionResult[]
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddb_ctx_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), alg.device, ctx))
    upload(Ym) = check(ccall((:ddb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Int64),
        ctx[], X, Ym, size(Ym, 1), m))
    gram() = check(ccall((:ddb_gram, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], C_NULL))
    setview(a, b) = check(ccall((:ddb_set_sample_range, LIB), Cint, (Ptr{Cvoid}, Int64, Int64), ctx[], a, b))
    try
        GC.@preserve X Y Yd E lambdas begin
            check(ccall((:ddb_set_library_monomial, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx[], n, k, E))
            upload(Yd)
            if unit
                # fit(UnitRangeTransform, Theta, dims = 2) on ALL samples; an infinite scale (constant row) becomes 1
                mx = zeros(k)
                check(ccall((:ddb_term_minmax, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx[], umin, mx))
                udiv .= [r > 0.0 ? r : 1.0 for r in (mx .- umin)]
            elseif zs
                # fit(ZScoreTransform, Theta, dims = 2, center = false): two-pass standard deviation of every row
                gram()
                sums = zeros(k)
                c = findfirst(j -> all(iszero, view_col(E, j)), 1:k)
                if c !== nothing               # row of the constant term = sum over the samples of every term
                    G = zeros(k, k)
                    check(ccall((:ddb_gram_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        ctx[], G, C_NULL, C_NULL))
                    sums .= G[:, c]
                else                           # one more pass with a row of ones as the target: R[:, 1] = Theta * 1
                    upload(ones(1, m)); gram()
                    R1 = zeros(k, 1)
                    check(ccall((:ddb_gram_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        ctx[], C_NULL, R1, C_NULL))
                    sums .= R1[:, 1]
                    upload(Yd)
                end
                center = sums ./ m; css = zeros(k)
                check(ccall((:ddb_term_css, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx[], center, css))
                sigma .= sqrt.(max.(css, 0.0) ./ (m - 1))
                sigma[sigma .== 0.0] .= 1.0    # constants are scaled with 1 (data_processing.jl:75-80)
            end
            batchsize = dp.batchsize <= 0 ? m_train : dp.batchsize
            if batchsize < m_train             # DataLoader(shuffle = true, rng = rng) (data_processing.jl:38)
                perm = Int64.(vcat(Random.randperm(dp.rng, m_train) .- 1, m_train:(m - 1)))
                GC.@preserve perm check(ccall((:ddb_permute_samples, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx[], perm))
            end
            starts = collect(0:batchsize:(m_train - 1))
            (!dp.partial && m_train % batchsize != 0) && pop!(starts)
            for b0 in starts
                b1 = min(b0 + batchsize, m_train); mb = b1 - b0
                setview(b0, b1); gram()
                zs && check(ccall((:ddb_gram_scale_terms, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], sigma))
                if unit
                    R = zeros(k, ntd); ymean = zeros(ntd); ycss = zeros(ntd)
                    check(ccall((:ddb_gram_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        ctx[], C_NULL, R, C_NULL))
                    check(ccall((:ddb_target_stats, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
                        ctx[], mb, ymean, ycss))
                    tsum = R[:, ntd]; ysum = ymean .* mb
                    check(ccall((:ddb_gram_affine_terms, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                        ctx[], umin, udiv, tsum, ysum, mb))
                end
                coef = zeros(ntd, k); lam = zeros(ntd); iters = zeros(Int32, ntd); rssv = zeros(ntd)
                dof = zeros(Int32, ntd); flags = zeros(Int32, ntd)
                check(ccall((:ddb_sweep, LIB), Cint, (Ptr{Cvoid}, Ptr{SweepParams}, Ptr{Float64}, Cint, Int64),
                    ctx[], prm, lambdas, L, mb))
                check(ccall((:ddb_rss, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], C_NULL))
                check(ccall((:ddb_select, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{UInt32},
                     Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                    ctx[], coef, lam, iters, rssv, dof, C_NULL, C_NULL, C_NULL, flags))
                Cs = coef[1:n_t, :]            # in the transformed space; the auxiliary ones target is dropped
                Cu = zs ? Cs ./ sigma' : (unit ? Cs ./ udiv' : Cs)
                testerror = nothing
                if m_train < m
                    setview(m_train, m)
                    r = zeros(ntd)
                    if unit                    # the transformed-space model on the transformed test rows, on the raw data
                        off = vcat(-(Cu * umin), 0.0); Cfull = vcat(Cu, zeros(1, k))
                        check(ccall((:ddb_residual_affine, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                            ctx[], Cfull, off, r))
                    else
                        check(ccall((:ddb_residual, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx[], Cu, r))
                    end
                    testerror = sum(r[1:n_t])
                end
                retcode = any(!iszero, flags[1:n_t] .& Int32(1)) ? DDReturnCode(2) : DDReturnCode(1)
                push!(results, DataDrivenSparse.SparseRegressionResult(Cs, sum(abs.(Cs) .> 0.0), lam[1:n_t],
                    Int.(iters[1:n_t]), testerror, sum(rssv[1:n_t]), retcode))
            end
        end
    finally
        ccall((:ddb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), ctx[])
    end
    # lib/DataDrivenSparse/src/commonsolve.jl:13-24
    sort!(results, by = DataDrivenSparse.l2error)
    best = first(results)
    C = copy(DataDrivenSparse.get_coefficients(best))
    C = unit ? (C .- umin') ./ udiv' : (zs ? C ./ sigma' : C)   # apply_transform on the transposed coefficients, as written
    new_basis = __construct_basis(C, basis, problem, options)
    return DataDrivenSolution(new_basis, problem, alg, results, ps, best.retcode)
end
view_col(E, j) = @view E[:, j]

# ------------------------------------------------------------------------------------------------------
# DataDrivenDMD: `B200DMD <: AbstractKoopmanAlgorithm` implementing the `alg(X, Y) -> (K::Eigen, B)` contract of
# lib/DataDrivenDMD/src/algorithms.jl:80-83,147-158.  The snapshot contractions P = X X', Q = Y X' and the
# DMDPINV operator K = Q P^-1 come from one `ddb_dmd_solve` call (the caller passes the snapshot matrix whose
# columns 1:m are X and 2:m+1 are Y, as `get_fit_targets` builds them for a discrete problem, solve.jl:36-65);
# the exact-DMD factors follow from the Gram identities  P = U S^2 U',  B = Y V S^-1 = Q U S^-2,  Atilde = U'B.
# (Requires `using DataDrivenDMD, LinearAlgebra` in the host session.)
# ------------------------------------------------------------------------------------------------------
struct B200DMD{T <: Real}
    truncation::T      # 0 => DMDPINV through the normal equations, otherwise DMDSVD(truncation)
    device::Int
end
B200DMD() = B200DMD(0.0, 0)

function dmd_blocks(alg::B200DMD, x::Matrix{Float64}; want_K::Bool = true)
    n, mp1 = size(x)
    P = zeros(n, n); Q = zeros(n, n); K = want_K ? zeros(n, n) : zeros(0, 0)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddb_ctx_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), alg.device, ctx))
    try
        GC.@preserve x P Q K begin
            check(ccall((:ddb_dmd_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx[], x, n, mp1, P, Q, want_K ? pointer(K) : Ptr{Float64}(C_NULL)))
        end
    finally
        ccall((:ddb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), ctx[])
    end
    return P, Q, K
end

# (alg::B200DMD)(x): x = [X Y[:, end]] -- one resident copy serves X and Y.  The eigen-decompositions are dense
# n x n LAPACK work on the host here (LinearAlgebra); `ddb_dmd_eig_sym` / `ddb_dmd_eig` / `ddb_dmd_residual` /
# `ddb_gemm_nt` offer the same steps on the device for state counts where that matters (see INTEGRATION.md).
function (alg::B200DMD)(x::AbstractMatrix)
    xs = Matrix{Float64}(x)
    if alg.truncation == 0
        _, _, K = dmd_blocks(alg, xs)
        return (LinearAlgebra.eigen(K), zeros(0, 0))
    end
    P, Q, _ = dmd_blocks(alg, xs; want_K = false)
    F = LinearAlgebra.eigen(LinearAlgebra.Symmetric(P))
    s2 = reverse(F.values); U = reverse(F.vectors, dims = 2)
    keep = s2 .> (min(alg.truncation, 1.0)^2 * s2[1])
    U = U[:, keep]; s2 = s2[keep]
    B = (Q * U) ./ s2'
    λ, ω = LinearAlgebra.eigen(U' * B)
    return (LinearAlgebra.Eigen(λ, B * ω), zeros(0, 0))
end

export B200Sparse, B200DMD
end # module
