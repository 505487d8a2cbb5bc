# OptDynB200.jl — Julia shim over liboptdyn_b200.so (include/optdyn_b200.h).
#
# Keeps the reference's API (src/dynamics.jl:1-145, src/gradient_bundle.jl:15-147, src/models/rocket/dynamics.jl:13-269) AND the slice
# of RoboDojo's surface the examples touch directly (examples/hopper.jl:3,14,38-50,63,97-99), so that every script under examples/
# runs with ONE changed line:
#
#     -using OptimizationDynamics
#     +include("julia/OptDynB200.jl"); using .OptDynB200; const OptimizationDynamics = OptDynB200
#
# (the following `const iLQR = OptimizationDynamics.IterativeLQR` / `const RoboDojo = OptimizationDynamics.RoboDojo` lines of the
# examples then resolve to the submodules below).  What is kept, by name and argument order:
#   * model singletons `acrobot_impact`, `acrobot_nominal`, `cartpole_friction`, `cartpole_frictionless`, `planarpush`, `rocket`,
#     `RoboDojo.hopper` (fields nq, nu, nw, friction, and the constants the examples read: foot_radius, mass_body, gravity, length);
#   * the generated-code handles `r_*_func`, `rz_*_func`, `rθ_*_func`, `RoboDojo.residual_expr(model)` … — placeholders (`:(nothing)`, so
#     `eval(...)` works): the residual code is compiled into the library (csrc/gen/) and the constructors ignore them;
#   * `ImplicitDynamics(model, h, r, rz, rθ; T, r_tol, κ_eval_tol, κ_grad_tol, no_impact, no_friction, n, m, d, nc, nb, info)` with
#     fields `eval_sim`, `grad_sim`, `info`, `idx_q1`, `idx_q2`, `idx_u1`;  `sim.h`, `sim.model.nq`,
#     `sim.grad.∂q3∂q1[1]`, `∂q3∂q2[1]`, `∂q3∂u1[1]`;  `RoboDojo.step!(sim, q2, v1, u1, t)`;
#   * `f`, `fx`, `fu`, `state_to_configuration`; `GradientBundle(model; N, ϵ)`, `fx_gb`, `fu_gb`; `RocketInfo(rocket, u_max, h, r…)`,
#     `f/fx/fu_rocket[_proj]`, `soc_projection[_gradient]`;
#   * `model.friction .= μ` after construction (examples/cartpole.jl:21) is observed at the next call;
#   * `Visualizer`, `render`, `visualize!` are no-ops (MeshCat rendering is outside the hot path, SURVEY.md §2 row 15).
#
# NOTE: this file cannot be executed in the build image (no Julia toolchain); it is the binding a maintainer adds.  The same
# calls are exercised from Python/ctypes by tests/test_gpu_parity.py (incl. a transliteration of examples/hopper.jl:52-160).
module OptDynB200

import IterativeLQR                       # same dependency as the reference (Project.toml:13): `OptimizationDynamics.IterativeLQR`

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

# ---- model singletons (reference src/models/*/model.jl; RoboDojo.hopper) -----------------------------------------------------------
struct Model
    name::Symbol
    nq::Int
    nu::Int
    nw::Int
    nc::Int
    friction::Vector{Float64}                 # mutable contents, like the reference's model.friction
    consts::Dict{Symbol,Float64}
end
function Base.getproperty(m::Model, s::Symbol)
    s in fieldnames(Model) && return getfield(m, s)
    c = getfield(m, :consts)
    haskey(c, s) || error("model $(getfield(m, :name)) has no field $s")
    return c[s]
end
const acrobot_impact = Model(:acrobot_impact, 2, 1, 0, 2, Float64[], Dict{Symbol,Float64}())
const acrobot_nominal = Model(:acrobot_nominal, 2, 1, 0, 0, Float64[], Dict{Symbol,Float64}())
const cartpole_friction = Model(:cartpole_friction, 2, 1, 0, 2, [0.1, 0.1], Dict(:mc => 1.0, :mp => 0.2, :l => 0.5, :g => 9.81))
const cartpole_frictionless = Model(:cartpole_frictionless, 2, 1, 0, 2, Float64[], Dict(:mc => 1.0, :mp => 0.2, :l => 0.5, :g => 9.81))
const planarpush = Model(:planar_push, 5, 2, 0, 5, Float64[], Dict{Symbol,Float64}())
const hopper = Model(:hopper, 4, 2, 0, 4, [0.5, 0.5], Dict(:mass_body => 3.0, :mass_foot => 1.0, :inertia_body => 0.75, :gravity => 9.81,
                     :body_radius => 0.1, :foot_radius => 0.05, :leg_len_max => 1.0, :leg_len_min => 0.25))
const rocket = Model(:rocket, 12, 3, 0, 0, Float64[], Dict(:mass => 1.0, :length => 1.0))

# generated-code handles of the reference (src/OptimizationDynamics.jl:75-88): placeholders, `eval(r_pp_func)` → nothing
for name in (:r_acrobot_impact_func, :rz_acrobot_impact_func, :rθ_acrobot_impact_func, :r_acrobot_nominal_func, :rz_acrobot_nominal_func,
             :rθ_acrobot_nominal_func, :r_cartpole_friction_func, :rz_cartpole_friction_func, :rθ_cartpole_friction_func,
             :r_cartpole_frictionless_func, :rz_cartpole_frictionless_func, :rθ_cartpole_frictionless_func, :r_pp_func, :rz_pp_func,
             :rθ_pp_func, :r_rocket_func, :rz_rocket_func, :rθ_rocket_func, :r_proj_func, :rz_proj_func, :rθ_proj_func)
    @eval const $name = :(nothing)
    @eval export $name
end

# ---- simulators -------------------------------------------------------------------------------------------------------------------
# One device handle per ImplicitDynamics, shared by its two simulator proxies; re-created when model.friction was mutated.
mutable struct HandleBox
    handle::Ptr{Cvoid}
    model::Model
    h::Float64
    opts::ODOptions
    device::Int
    friction_seen::Vector{Float64}
end
function create!(box::HandleBox)
    fr = box.model.friction
    hd = ccall((:od_create, LIB), Ptr{Cvoid}, (Cint, Cdouble, Ref{ODOptions}, Ptr{Cdouble}, Cint, Cint),
               MODEL_ID[box.model.name], box.h, Ref(box.opts), fr, length(fr), box.device)
    hd == C_NULL && error("optdyn_b200: " * last_error())
    box.handle != C_NULL && ccall((:od_destroy, LIB), Cvoid, (Ptr{Cvoid},), box.handle)
    box.handle = hd
    box.friction_seen = copy(fr)
    return box
end
function handle(box::HandleBox)
    box.model.friction != box.friction_seen && create!(box)
    return box.handle
end

struct SimulatorGrad                       # sim.grad as the examples read it (T = 1 ⇒ 1-element vectors; src/dynamics.jl:39-46)
    ∂q3∂q1::Vector{Matrix{Float64}}
    ∂q3∂q2::Vector{Matrix{Float64}}
    ∂q3∂u1::Vector{Matrix{Float64}}
end
mutable struct Simulator                   # what model.eval_sim / model.grad_sim are to the reference's callers
    box::HandleBox
    model::Model
    h::Float64
    diff_sol::Bool
    grad::SimulatorGrad
    q3::Vector{Float64}
    status::Vector{Int32}
end
Simulator(box::HandleBox, diff_sol::Bool) = Simulator(box, box.model, box.h, diff_sol,
    SimulatorGrad([zeros(box.model.nq, box.model.nq)], [zeros(box.model.nq, box.model.nq)], [zeros(box.model.nq, box.model.nu)]),
    zeros(box.model.nq), Int32[0])

# RoboDojo.step!(sim, q, v, u, t) — examples/hopper.jl:63,89,112,133,157; src/dynamics.jl:88,103,123
function step!(sim::Simulator, q, v, u, t=1)
    qv = convert(Vector{Float64}, q); vv = convert(Vector{Float64}, v); uv = convert(Vector{Float64}, u)
    if sim.diff_sol
        check(ccall((:od_sim_step_batch, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                    handle(sim.box), 1, 1, qv, vv, uv, sim.q3, sim.grad.∂q3∂q1[1], sim.grad.∂q3∂q2[1], sim.grad.∂q3∂u1[1], sim.status))
    else
        check(ccall((:od_sim_step_batch, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                    handle(sim.box), 1, 0, qv, vv, uv, sim.q3, C_NULL, C_NULL, C_NULL, sim.status))
    end
    return sim.q3
end

mutable struct ImplicitDynamics{I}
    n::Int
    m::Int
    d::Int
    eval_sim::Simulator
    grad_sim::Simulator
    nq::Int
    nu::Int
    h::Float64
    idx_q1::Vector{Int}
    idx_q2::Vector{Int}
    idx_u1::Vector{Int}
    info::I
    box::HandleBox
    # fx and fu share one solve (the reference solves twice, src/dynamics.jl:103,123)
    memo_x::Vector{Float64}
    memo_u::Vector{Float64}
    memo_friction::Vector{Float64}
    dq1::Matrix{Float64}
    dq2::Matrix{Float64}
    du1::Matrix{Float64}
    q3::Vector{Float64}
    status::Vector{Int32}
end

# src/dynamics.jl:51-79 — same positional and keyword arguments; r / rz / rθ are ignored
function ImplicitDynamics(model::Model, h, r_func=nothing, rz_func=nothing, rθ_func=nothing;
        T=1, r_tol=1.0e-8, κ_eval_tol=1.0e-6, κ_grad_tol=1.0e-6, no_impact=false, no_friction=false,
        n=2 * model.nq, m=model.nu, d=model.nw, nc=model.nc, nb=model.nc, info=nothing, device=0)
    opts = ODOptions(r_tol, κ_eval_tol, κ_grad_tol, 0.5, 100, 25)               # src/dynamics.jl:25-33
    box = create!(HandleBox(C_NULL, model, Float64(h), opts, device, Float64[]))
    nq = model.nq; nu = model.nu
    obj = ImplicitDynamics(n, m, d, Simulator(box, false), Simulator(box, true), nq, nu, Float64(h),
                           collect(1:nq), collect(nq .+ (1:nq)), collect(1:nu), info, box,
                           Float64[], Float64[], Float64[], zeros(nq, nq), zeros(nq, nq), zeros(nq, nu), zeros(nq), Int32[0])
    finalizer(o -> (o.box.handle != C_NULL && ccall((:od_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.box.handle); o.box.handle = C_NULL), obj)
    return obj
end

# f — src/dynamics.jl:81-94
function f(d, model::ImplicitDynamics, x, u, w)
    q1 = x[model.idx_q1]; q2 = x[model.idx_q2]; u1 = u[model.idx_u1]
    check(ccall((:od_step_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                handle(model.box), 1, q1, q2, u1, model.q3, model.status))
    d[model.idx_q1] .= q2
    d[model.idx_q2] .= model.q3
    return d
end

function _grad!(model::ImplicitDynamics, x, u)
    fr = model.box.model.friction
    if model.memo_x != x || model.memo_u != u || model.memo_friction != fr
        q1 = x[model.idx_q1]; q2 = x[model.idx_q2]; u1 = u[model.idx_u1]
        # blocks come back column-major = Julia layout: no transpose
        check(ccall((:od_step_grad_batch, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                    handle(model.box), 1, q1, q2, u1, C_NULL, model.dq1, model.dq2, model.du1, model.status))
        model.memo_x = copy(x); model.memo_u = copy(u); model.memo_friction = copy(fr)
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
                handle(model.box), B, q1, q2, u1, q3, dq1, dq2, du1, status))
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
                handle(model.box), R, T, x1, ū, 0, x̄ === nothing ? C_NULL : x̄, Kc, k === nothing ? C_NULL : k, α === nothing ? C_NULL : α, X, U, status))
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
                handle(model.box), NT, S + 1, jac, lx, lu, lxx, luu, lux === nothing ? C_NULL : lux, reg, K, k, ΔV, status))
    return K, k, ΔV, status
end

# state_to_configuration — same contract as src/dynamics.jl:131-145: [x₁[1:nq], x₁[nq+1:2nq], x₂[nq+1:2nq], …]
function state_to_configuration(x::Vector{Vector{T}}) where T
    nq = length(x[1]) ÷ 2
    q = [x[1][1:nq]]
    append!(q, [xt[nq .+ (1:nq)] for xt in x])
    return q
end

# ---- gradient bundle — src/gradient_bundle.jl:15-147 ----------------------------------------------------------------------------
struct GradientBundle
    η::Matrix{Float64}      # (2nq+nu) × N, columns = perturbations
    dz::Matrix{Float64}
end
# GradientBundle(model; N, ϵ) — src/gradient_bundle.jl:26 (buffers sized from model.nq: the reference's module-global `nq` bug,
# src/gradient_bundle.jl:79-80, is not reproduced)
function GradientBundle(model::Model; N=100, ϵ=1.0e-4)
    nq = model.nq; nz = 2nq + model.nu
    η = zeros(nz, N)
    for i = 1:N
        η[rand(1:nz), i] = ϵ * randn()          # src/gradient_bundle.jl:49-54
    end
    GradientBundle(η, zeros(nq, nz))
end
# gradient!(sim, gb, q1, q2, u1) — src/gradient_bundle.jl:87-104; also accepts the ImplicitDynamics that owns the simulator
function gradient!(sim::Simulator, gb::GradientBundle, q1, q2, u1)
    status = Int32[0]
    check(ccall((:od_bundle_batch, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                handle(sim.box), 1, size(gb.η, 2), gb.η, convert(Vector{Float64}, q1), convert(Vector{Float64}, q2), convert(Vector{Float64}, u1), gb.dz, status))
    return gb.dz
end
gradient!(model::ImplicitDynamics, gb::GradientBundle, q1, q2, u1) = gradient!(model.eval_sim, gb, q1, q2, u1)
function fx_gb(dx, model::ImplicitDynamics, x, u, w)
    nq = model.nq
    for i = 1:nq
        dx[model.idx_q1[i], model.idx_q2[i]] = 1.0
    end
    dz = gradient!(model.eval_sim, model.info, x[model.idx_q1], x[model.idx_q2], u[model.idx_u1])
    dx[model.idx_q2, model.idx_q1] = dz[:, 1:nq]
    dx[model.idx_q2, model.idx_q2] = dz[:, nq .+ (1:nq)]
    return dx
end
function fu_gb(du, model::ImplicitDynamics, x, u, w)
    nq = model.nq
    dz = gradient!(model.eval_sim, model.info, x[model.idx_q1], x[model.idx_q2], u[model.idx_u1])
    du[model.idx_q2, :] = dz[:, 2nq .+ (1:model.nu)]
    return du
end

# ---- rocket — src/models/rocket/dynamics.jl:13-269 ---------------------------------------------------------------------------------
mutable struct RocketInfo
    handle::Ptr{Cvoid}
    h::Float64
    u_max::Float64
    # scratch: fx / fu of the reference each return one block; the other outputs of the shared solve land here (no allocation per call)
    y::Vector{Float64}
    dx::Matrix{Float64}
    du::Matrix{Float64}
    status::Vector{Int32}
end
function RocketInfo(rocket, u_max, h, generated...; device=0)
    hd = ccall((:od_create, LIB), Ptr{Cvoid}, (Cint, Cdouble, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint), MODEL_ID[:rocket], h, C_NULL, [Float64(u_max)], 1, device)
    hd == C_NULL && error("optdyn_b200: " * last_error())
    obj = RocketInfo(hd, h, u_max, zeros(12), zeros(12, 12), zeros(12, 3), Int32[0])
    finalizer(o -> ccall((:od_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle), obj)
    return obj
end
function _rocket(info::RocketInfo, x, u, proj::Bool, y, dx, du)
    check(ccall((:od_rocket_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                info.handle, 1, convert(Vector{Float64}, x), convert(Vector{Float64}, u), proj ? 1 : 0, y, dx, du, info.status))
end
f_rocket(d, info::RocketInfo, x, u, w) = (_rocket(info, x, u, false, d, C_NULL, C_NULL); d)
fx_rocket(dx, info::RocketInfo, x, u, w) = (_rocket(info, x, u, false, info.y, dx, info.du); dx)
fu_rocket(du, info::RocketInfo, x, u, w) = (_rocket(info, x, u, false, info.y, info.dx, du); du)
f_rocket_proj(d, info::RocketInfo, x, u, w) = (_rocket(info, x, u, true, d, C_NULL, C_NULL); d)
fx_rocket_proj(dx, info::RocketInfo, x, u, w) = (_rocket(info, x, u, true, info.y, dx, info.du); dx)
fu_rocket_proj(du, info::RocketInfo, x, u, w) = (_rocket(info, x, u, true, info.y, info.dx, du); du)
function soc_projection(x, info::RocketInfo)
    up = zeros(3)
    check(ccall((:od_rocket_projection_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}), info.handle, 1, convert(Vector{Float64}, x), up, C_NULL, info.status))
    return up
end
function soc_projection_gradient(x, info::RocketInfo)
    up = zeros(3); dup = zeros(3, 3)
    check(ccall((:od_rocket_projection_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}), info.handle, 1, convert(Vector{Float64}, x), up, dup, info.status))
    return dup
end

# ---- visualisation: out of scope (SURVEY.md §2 row 15); no-ops so that the examples run unchanged ---------------------------------------
struct Visualizer end
render(vis) = nothing
visualize!(args...; kwargs...) = nothing

# ---- the slice of RoboDojo the examples use (examples/hopper.jl:3,14,39-41,63,273) ------------------------------------------------------
residual_expr(model::Model) = :(nothing)
jacobian_var_expr(model::Model) = :(nothing)
jacobian_data_expr(model::Model) = :(nothing)
friction_coefficients(model::Model) = model.friction
module RoboDojo
    import ..OptDynB200: hopper, step!, residual_expr, jacobian_var_expr, jacobian_data_expr, friction_coefficients, visualize!, Simulator
end

export ImplicitDynamics, Simulator, f, fx, fu, step!, step_grad_batch, rollout, rollout_batch, backward_pass_batch, state_to_configuration,
       GradientBundle, gradient!, fx_gb, fu_gb, RocketInfo, f_rocket, fx_rocket, fu_rocket, f_rocket_proj, fx_rocket_proj, fu_rocket_proj,
       soc_projection, soc_projection_gradient, acrobot_impact, acrobot_nominal, cartpole_friction, cartpole_frictionless, planarpush, rocket,
       Visualizer, render, visualize!, RoboDojo, IterativeLQR
end # module
