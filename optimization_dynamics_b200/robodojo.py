"""The slice of RoboDojo.jl's surface that the reference's examples touch directly (reference examples/hopper.jl:3,14,38-50,63):

    RoboDojo.hopper                                   the model singleton (nq, nu, foot_radius, mass_body, gravity, …)
    RoboDojo.residual_expr / jacobian_var_expr / jacobian_data_expr(model)
                                                      generated-code handles: placeholders here — the residual code is compiled
                                                      into liboptdyn_b200.so (csrc/gen/), ImplicitDynamics ignores them
    RoboDojo.step!(sim, q, v, u, t)                   → `step(sim, q, v, u, t)`; sim = im_dyn.eval_sim / im_dyn.grad_sim

so that `f1 / f1u / ft / ftx / ftu` of the hopper example transliterate line by line (tests/test_gpu_parity.py)."""
from .dynamics import hopper  # noqa: F401


def residual_expr(model):
    return None


def jacobian_var_expr(model):
    return None


def jacobian_data_expr(model):
    return None


def step(sim, q, v, u, t=1):
    """RoboDojo.step!(sim, q, v, u, t): q1 = q − h·v, initialize_z!, interior-point solve, q3 (and sim.grad when diff_sol)."""
    return sim.step(q, v, u, t)
