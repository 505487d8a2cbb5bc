# OptDynB200.jl — Julia shim over liboptdyn_b200.so (include/optdyn_b200.h).
#
# Keeps the reference's API (src/dynamics.jl:1-145, src/gradient_bundle.jl:15-147, src/models/rocket/dynamics.jl:13-269):
#   ImplicitDynamics(model, h, r, rz, rθ; r_tol, κ_eval_tol, κ_grad_tol, …), f, fx, fu, fx_gb, fu_gb, RocketInfo, f_rocket_proj, …
# so the examples run unchanged: `iLQR.Dynamics((d,x,u,w)->f(d,im_dyn,x,u,w), …)` (examples/cartpole.jl:34-37).
# The three generated-function arguments (r, rz, rθ) are accepted and ignored — the residual code lives in the library.
#
# NOTE: this file cannot be executed in the build image (no Julia toolchain); it is the binding a maintainer adds.  The same
# calls are exercised from Python/ctypes by tests/test_gpu_parity.py.
module OptDynB200

const LIB = get(ENV, "OPTDYN_B200_LIB", joinpath(@__DIR__, "..", "optimization_dynamics_b200", "liboptdyn_b200.so"))

const MODEL_ID = Dict(:acrobot_impact => 0, :acrobot_nominal => 1, :cartpole_friction => 2, :cartpole_frictionless => 3,
                      :planar_push => 4, :hopper => 5, :rocket => 6)

struct ODOptions          # od_options
    r_tol::Cdouble
    kappa_eval_tol::Cdouble
    kappa_grad_tol::Cdouble
    ls_scale::Cdouble
    max_iter::Int32
    max_ls::Int32
end

last_error() = unsafe_string(ccall((:od_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("optdyn_b200: " * last_error())

mutable struct ImplicitDynamics
    handle::Ptr{Cvoid}
    nq::Int
    nu::Int
    h::Float64
    idx_q1::Vector{Int}
    idx_q2::Vector{Int}
    idx_u1::Vector{Int}
    info::Any
    # fx and fu share one solve (the reference solves twice, src/dynamics.jl:103,123)
    memo_x::Vector{Float64}
    memo_u::Vector{Float64}
    dq1::Matrix{Float64}
    dq2::Matrix{Float64}
    du1::Matrix{Float64}
    q3::Vector{Float64}
    status::Vector{Int32}
end

function ImplicitDynamics(model::Symbol, h, r_func=nothing, rz_func=nothing, rθ_func=nothing;
        T=1, r_tol=1.0e-8, κ_eval_tol=1.0e-6, κ_grad_tol=1.0e-6, no_impact=false, no_friction=false,
        n=nothing, m=nothing, d=nothing, nc=nothing, nb=nothing, info=nothing, friction=Float64[], device=0)
    id = MODEL_ID[model]
    nq = Ref{Cint}(0); nu = Ref{Cint}(0); nz = Ref{Cint}(0); nθ = Ref{Cint}(0)
    check(ccall((:od_model_dims, LIB), Cint, (Cint, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}), id, nq, nu, nz, nθ))
    opts = Ref(ODOptions(r_tol, κ_eval_tol, κ_grad_tol, 0.5, 100, 25))         # src/dynamics.jl:25-33
    hd = ccall((:od_create, LIB), Ptr{Cvoid}, (Cint, Cdouble, Ref{ODOptions}, Ptr{Cdouble}, Cint, Cint),
               id, h, opts, friction, length(friction), device)
    hd == C_NULL && error("optdyn_b200: " * last_error())
    obj = ImplicitDynamics(hd, nq[], nu[], h, collect(1:nq[]), collect(nq[] .+ (1:nq[])), collect(1:nu[]), info,
                           Float64[], Float64[], zeros(nq[], nq[]), zeros(nq[], nq[]), zeros(nq[], nu[]), zeros(nq[]), Int32[0])
    finalizer(o -> ccall((:od_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle), obj)
    return obj
end

# f — src/dynamics.jl:81-94
function f(d, model::ImplicitDynamics, x, u, w)
    q1 = x[model.idx_q1]; q2 = x[model.idx_q2]; u1 = u[model.idx_u1]
    check(ccall((:od_step_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                model.handle, 1, q1, q2, u1, model.q3, model.status))
    d[model.idx_q1] .= q2
    d[model.idx_q2] .= model.q3
    return d
end

function _grad!(model::ImplicitDynamics, x, u)
    if model.memo_x != x || model.memo_u != u
        q1 = x[model.idx_q1]; q2 = x[model.idx_q2]; u1 = u[model.idx_u1]
        # blocks come back column-major = Julia layout: no transpose
        check(ccall((:od_step_grad_batch, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                    model.handle, 1, q1, q2, u1, C_NULL, model.dq1, model.dq2, model.du1, model.status))
        model.memo_x = copy(x); model.memo_u = copy(u)
    end
    return nothing
end

# fx — src/dynamics.jl:96-114 (only the three blocks are written, like the reference)
function fx(dx, model::ImplicitDynamics, x, u, w)
    _grad!(model, x, u)
    for i = 1:model.nq
        dx[model.idx_q1[i], model.idx_q2[i]] = 1.0
    end
    dx[model.idx_q2, model.idx_q1] .= model.dq1
    dx[model.idx_q2, model.idx_q2] .= model.dq2
    return dx
end

# fu — src/dynamics.jl:116-128
function fu(du, model::ImplicitDynamics, x, u, w)
    _grad!(model, x, u)
    du[model.idx_q2, :] .= model.du1
    return du
end

# Batched derivative sweep: all timesteps (× samples × rollouts) in one launch.  X is 2nq×B, U is nu×B (columns = problems);
# returns q3 (nq×B) and ∂q3∂q1, ∂q3∂q2 (nq×nq×B), ∂q3∂u1 (nq×nu×B) — the arrays the Riccati backward pass consumes.
function step_grad_batch(model::ImplicitDynamics, X::Matrix{Float64}, U::Matrix{Float64})
    B = size(X, 2); nq = model.nq; nu = model.nu
    q1 = X[model.idx_q1, :]; q2 = X[model.idx_q2, :]; u1 = U[model.idx_u1, :]      # nq×B column-major == B rows of nq for the C ABI
    q3 = zeros(nq, B); dq1 = zeros(nq, nq, B); dq2 = zeros(nq, nq, B); du1 = zeros(nq, nu, B); status = zeros(Int32, B)
    check(ccall((:od_step_grad_batch, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                model.handle, B, q1, q2, u1, q3, dq1, dq2, du1, status))
    return q3, dq1, dq2, du1, status
end

# rollout(model, x1, ū) — iLQR.rollout (examples/cartpole.jl:79): T−1 sequential calls of f in ONE launch.  ū is nu×(T−1).
# With a policy (x̄ 2nq×T, K as a vector of nu×2nq matrices, k nu×(T−1)) and step sizes α it is IterativeLQR's forward pass for all
# line-search candidates at once: u[t] = ū[t] + α k[t] + K[t](x[t] − x̄[t]).  Returns X (2nq×T×R), U (nu×(T−1)×R), status ((T−1)×R).
function rollout_batch(model::ImplicitDynamics, x1::Matrix{Float64}, ū::Matrix{Float64};
                       x̄=nothing, K=nothing, k=nothing, α=nothing)
    R = size(x1, 2); S = size(ū, 2); T = S + 1; nx = 2 * model.nq; nu = model.nu
    X = zeros(nx, T, R); U = zeros(nu, S, R); status = zeros(Int32, S, R)
    Kc = K === nothing ? C_NULL : permutedims(cat(K...; dims=3), (2, 1, 3))      # [t][control][state] row-major = state-fastest
    check(ccall((:od_rollout_batch, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                model.handle, R, T, x1, ū, 0, x̄ === nothing ? C_NULL : x̄, Kc, k === nothing ? C_NULL : k, α === nothing ? C_NULL : α, X, U, status))
    return X, U, status
end
rollout(model::ImplicitDynamics, x1::Vector{Float64}, ū::Vector{Vector{Float64}}) =
    (X = rollout_batch(model, reshape(x1, :, 1), reduce(hcat, ū))[1]; [X[:, t, 1] for t = 1:size(X, 2)])

# backward_pass_batch — IterativeLQR's Riccati recursion for NT trajectories at once, on the packed rows of the derivative sweep
# (jac: (nq + nq(2nq+nu))×(T−1)×NT).  lx n×T×NT, lu m×(T−1)×NT, lxx n×n×T×NT, luu m×m×(T−1)×NT (symmetric, so Julia's column-major
# blocks are the row-major blocks the ABI expects), lux optional as n×m×(T−1)×NT (= row-major m×n).  Returns K (n×m×(T−1)×NT, i.e.
# K[:, :, t, a]' is the nu×2nq gain), k (m×(T−1)×NT), ΔV (2×NT), status.
function backward_pass_batch(model::ImplicitDynamics, jac, lx, lu, lxx, luu; lux=nothing, reg=0.0)
    NT = size(jac, 3); S = size(jac, 2); n = 2 * model.nq; m = model.nu
    K = zeros(n, m, S, NT); k = zeros(m, S, NT); ΔV = zeros(2, NT); status = zeros(Int32, NT)
    check(ccall((:od_riccati_batch, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                model.handle, NT, S + 1, jac, lx, lu, lxx, luu, lux === nothing ? C_NULL : lux, reg, K, k, ΔV, status))
    return K, k, ΔV, status
end

# state_to_configuration — src/dynamics.jl:131-145
function state_to_configuration(x::Vector{Vector{T}}) where T
    nq = convert(Int, floor(length(x[1]) / 2))
    q = Vector{T}[]
    for t = 1:length(x)
        t == 1 && push!(q, x[t][1:nq])
        push!(q, x[t][nq .+ (1:nq)])
    end
    return q
end

# ---- gradient bundle — src/gradient_bundle.jl:15-147 ----------------------------------------------------------------------------
struct GradientBundle
    η::Matrix{Float64}      # (2nq+nu) × N, columns = perturbations
    dz::Matrix{Float64}
end
function GradientBundle(nq::Int, nu::Int; N=100, ϵ=1.0e-4)
    nz = 2nq + nu
    η = zeros(nz, N)
    for i = 1:N
        η[rand(1:nz), i] = ϵ * randn()          # src/gradient_bundle.jl:49-54
    end
    GradientBundle(η, zeros(nq, nz))
end
function gradient!(model::ImplicitDynamics, gb::GradientBundle, q1, q2, u1)
    status = Int32[0]
    check(ccall((:od_bundle_batch, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                model.handle, 1, size(gb.η, 2), gb.η, q1, q2, u1, gb.dz, status))
    return gb.dz
end
function fx_gb(dx, model::ImplicitDynamics, x, u, w)
    nq = model.nq
    for i = 1:nq
        dx[model.idx_q1[i], model.idx_q2[i]] = 1.0
    end
    dz = gradient!(model, model.info, x[model.idx_q1], x[model.idx_q2], u[model.idx_u1])
    dx[model.idx_q2, model.idx_q1] = dz[:, 1:nq]
    dx[model.idx_q2, model.idx_q2] = dz[:, nq .+ (1:nq)]
    return dx
end
function fu_gb(du, model::ImplicitDynamics, x, u, w)
    nq = model.nq
    dz = gradient!(model, model.info, x[model.idx_q1], x[model.idx_q2], u[model.idx_u1])
    du[model.idx_q2, :] = dz[:, 2nq .+ (1:model.nu)]
    return du
end

# ---- rocket — src/models/rocket/dynamics.jl:13-269 ---------------------------------------------------------------------------------
mutable struct RocketInfo
    handle::Ptr{Cvoid}
    h::Float64
    u_max::Float64
end
function RocketInfo(rocket, u_max, h, generated...; device=0)
    hd = ccall((:od_create, LIB), Ptr{Cvoid}, (Cint, Cdouble, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint), MODEL_ID[:rocket], h, C_NULL, [Float64(u_max)], 1, device)
    hd == C_NULL && error("optdyn_b200: " * last_error())
    obj = RocketInfo(hd, h, u_max)
    finalizer(o -> ccall((:od_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle), obj)
    return obj
end
function _rocket(info::RocketInfo, x, u, proj::Bool, y, dx, du)
    status = Int32[0]
    check(ccall((:od_rocket_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                info.handle, 1, x, u, proj ? 1 : 0, y, dx, du, status))
end
f_rocket(d, info::RocketInfo, x, u, w) = (_rocket(info, x, u, false, d, C_NULL, C_NULL); d)
fx_rocket(dx, info::RocketInfo, x, u, w) = (_rocket(info, x, u, false, zeros(12), dx, zeros(12, 3)); dx)
fu_rocket(du, info::RocketInfo, x, u, w) = (_rocket(info, x, u, false, zeros(12), zeros(12, 12), du); du)
f_rocket_proj(d, info::RocketInfo, x, u, w) = (_rocket(info, x, u, true, d, C_NULL, C_NULL); d)
fx_rocket_proj(dx, info::RocketInfo, x, u, w) = (_rocket(info, x, u, true, zeros(12), dx, zeros(12, 3)); dx)
fu_rocket_proj(du, info::RocketInfo, x, u, w) = (_rocket(info, x, u, true, zeros(12), zeros(12, 12), du); du)
function soc_projection(x, info::RocketInfo)
    up = zeros(3); status = Int32[0]
    check(ccall((:od_rocket_projection_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}), info.handle, 1, x, up, C_NULL, status))
    return up
end
function soc_projection_gradient(x, info::RocketInfo)
    up = zeros(3); dup = zeros(3, 3); status = Int32[0]
    check(ccall((:od_rocket_projection_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}), info.handle, 1, x, up, dup, status))
    return dup
end

export ImplicitDynamics, f, fx, fu, step_grad_batch, rollout, rollout_batch, backward_pass_batch, state_to_configuration, GradientBundle, gradient!, fx_gb, fu_gb,
       RocketInfo, f_rocket, fx_rocket, fu_rocket, f_rocket_proj, fx_rocket_proj, fu_rocket_proj, soc_projection, soc_projection_gradient
end # module
