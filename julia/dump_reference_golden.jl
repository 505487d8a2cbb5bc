# dump_reference_golden.jl — PINS PARITY: runs the UNMODIFIED reference (thowell/optimization_dynamics + RoboDojo.jl 0.1.x) on the
# committed golden inputs and writes what its own f / fx / fu return, so that tests/test_reference_golden.py can hold the oracle
# AND the CUDA path to the real Julia CPU path (1e-8 on q3, 1e-6 on the sensitivities — BASELINE.json north_star).
#
# The build image of this repository has no Julia, RoboDojo.jl is an un-vendored registry dependency of the reference
# (Project.toml:17,31), and there is no network: this script has therefore never been executed by the authors of this repo.
# Run it once on any machine with the reference installed:
#
#     julia --project=/path/to/optimization_dynamics julia/dump_reference_golden.jl
#
# and commit tests/golden/reference/*.csv.  Until those files exist the test skips with a loud reason and DESIGN.md §5 says
# "parity unpinned".
#
# Input  tests/golden/reference_inputs/<name>.csv : one row per problem, [q1 | q2 | u]   (rocket: [x | u]), written by
#        tests/golden/make_golden.py from the same seeded batches as tests/golden/<name>.npz.
# Output tests/golden/reference/<name>.csv        : one row per problem,
#        [q3 (nq) | ∂q3∂q1 (nq², column-major) | ∂q3∂q2 (nq²) | ∂q3∂u1 (nq·nu) | iterations eval | iterations grad]
#        (rocket: [y (12) | dx (144, column-major) | du (36)]).
# Calls made — exactly the reference's public path:
#     f(d, im_dyn, x, u, w)   src/dynamics.jl:81-94      fx(dx, …)  :96-114      fu(du, …)  :116-128
#     f_rocket[_proj] / fx_rocket[_proj] / fu_rocket[_proj]   src/models/rocket/dynamics.jl:101-163,215-269
using OptimizationDynamics
using DelimitedFiles
const RoboDojo = OptimizationDynamics.RoboDojo

const ROOT = normpath(joinpath(@__DIR__, ".."))
const IN = joinpath(ROOT, "tests", "golden", "reference_inputs")
const OUT = joinpath(ROOT, "tests", "golden", "reference")
mkpath(OUT)

iterations(sim) = try sim.ip.iterations catch; -1 end

function dump_contact(name, im_dyn, nq, nu)
    X = readdlm(joinpath(IN, name * ".csv"), ',', Float64)
    B = size(X, 1)
    res = zeros(B, nq + 2 * nq * nq + nq * nu + 2)
    w = zeros(0)
    for i = 1:B
        x = X[i, 1:2nq]; u = X[i, 2nq .+ (1:nu)]
        d = zeros(2nq); dx = zeros(2nq, 2nq); du = zeros(2nq, nu)
        f(d, im_dyn, x, u, w);   it_e = iterations(im_dyn.eval_sim)
        fx(dx, im_dyn, x, u, w); it_g = iterations(im_dyn.grad_sim)
        fu(du, im_dyn, x, u, w)
        res[i, :] = vcat(d[nq .+ (1:nq)], vec(dx[nq .+ (1:nq), 1:nq]), vec(dx[nq .+ (1:nq), nq .+ (1:nq)]), vec(du[nq .+ (1:nq), :]), it_e, it_g)
    end
    writedlm(joinpath(OUT, name * ".csv"), res, ',')
    println(name, ": ", B, " problems written")
end

# constructor calls = the reference's examples (examples/acrobot.jl:19-27, cartpole.jl:18-28, planar_push.jl:21-22, hopper.jl:38-43)
dump_contact("acrobot_impact",
    ImplicitDynamics(acrobot_impact, 0.05, eval(r_acrobot_impact_func), eval(rz_acrobot_impact_func), eval(rθ_acrobot_impact_func);
        r_tol=1.0e-8, κ_eval_tol=1.0e-4, κ_grad_tol=1.0e-3, no_friction=true), 2, 1)
dump_contact("acrobot_nominal",
    ImplicitDynamics(acrobot_nominal, 0.05, eval(r_acrobot_nominal_func), eval(rz_acrobot_nominal_func), eval(rθ_acrobot_nominal_func);
        r_tol=1.0e-8, κ_eval_tol=1.0e-4, κ_grad_tol=1.0e-3, no_friction=true), 2, 1)
im_cf = ImplicitDynamics(cartpole_friction, 0.05, eval(r_cartpole_friction_func), eval(rz_cartpole_friction_func), eval(rθ_cartpole_friction_func);
        r_tol=1.0e-8, κ_eval_tol=1.0e-4, κ_grad_tol=1.0e-3, no_impact=true)
cartpole_friction.friction .= [0.35; 0.35]                      # examples/cartpole.jl:21
dump_contact("cartpole_friction", im_cf, 2, 1)
dump_contact("cartpole_frictionless",
    ImplicitDynamics(cartpole_frictionless, 0.05, eval(r_cartpole_frictionless_func), eval(rz_cartpole_frictionless_func), eval(rθ_cartpole_frictionless_func);
        r_tol=1.0e-8, κ_eval_tol=1.0e-4, κ_grad_tol=1.0e-3, no_impact=true, no_friction=true), 2, 1)
dump_contact("planar_push",
    ImplicitDynamics(planarpush, 0.1, eval(r_pp_func), eval(rz_pp_func), eval(rθ_pp_func);
        r_tol=1.0e-8, κ_eval_tol=1.0e-4, κ_grad_tol=1.0e-2, nc=1, nb=9), 5, 2)
hopper = RoboDojo.hopper
dump_contact("hopper",
    ImplicitDynamics(hopper, 0.05, eval(RoboDojo.residual_expr(hopper)), eval(RoboDojo.jacobian_var_expr(hopper)), eval(RoboDojo.jacobian_data_expr(hopper));
        r_tol=1.0e-8, κ_eval_tol=1.0e-4, κ_grad_tol=1.0e-3, nc=4, nb=2), 4, 2)
# the hopper constants the oracle had to recollect (DESIGN.md §5) — printed so that a mismatch is visible at once
println("hopper constants: ", [(n, getfield(hopper, n)) for n in fieldnames(typeof(hopper)) if getfield(hopper, n) isa Number])
println("hopper friction:  ", RoboDojo.friction_coefficients(hopper))

# rocket (examples/rocket.jl:16-23)
info = RocketInfo(rocket, 12.5, 0.05,
    eval(r_rocket_func), eval(rz_rocket_func), eval(rθ_rocket_func),
    eval(r_proj_func), eval(rz_proj_func), eval(rθ_proj_func))
for (name, ff, ffx, ffu) in (("rocket", f_rocket, fx_rocket, fu_rocket), ("rocket_proj", f_rocket_proj, fx_rocket_proj, fu_rocket_proj))
    X = readdlm(joinpath(IN, "rocket.csv"), ',', Float64)
    B = size(X, 1)
    res = zeros(B, 12 + 144 + 36)
    for i = 1:B
        x = X[i, 1:12]; u = X[i, 13:15]
        d = zeros(12); dx = zeros(12, 12); du = zeros(12, 3)
        ff(d, info, x, u, zeros(0)); ffx(dx, info, x, u, zeros(0)); ffu(du, info, x, u, zeros(0))
        res[i, :] = vcat(d, vec(dx), vec(du))
    end
    writedlm(joinpath(OUT, name * ".csv"), res, ',')
    println(name, ": ", B, " problems written")
end
