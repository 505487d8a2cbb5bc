"""Synthetic (q1, q2, u) batches for every BASELINE.json config.  numpy only; seeded, reproducible.

The states come from the reference examples' initial conditions (examples/hopper.jl:178,270, examples/acrobot.jl:40-46,
examples/cartpole.jl:41-47, examples/planar_push.jl:46-53, examples/rocket.jl:44-55) spread out with seeded noise so that a
batch covers contact and no-contact regimes.
"""
import numpy as np

HOPPER = dict(mass_body=3.0, mass_foot=1.0, gravity=9.81, foot_radius=0.05, body_radius=0.1)


def hopper_batch(B, h=0.05, seed=0):
    """Half the batch with the foot on the ground (even indices), half in flight (odd indices)."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(-0.3, 0.3, B)
    r = rng.uniform(0.35, 0.9, B)
    x = rng.uniform(-0.5, 0.5, B)
    c = rng.uniform(0.0, 0.3, B)
    c[0::2] = 0.0
    z = HOPPER["foot_radius"] + r * np.cos(t) + c
    q2 = np.stack([x, z, t, r], axis=1)
    v = rng.normal(0.0, 0.5, (B, 4))
    q1 = q2 - h * v
    u = np.array([0.0, HOPPER["gravity"] * HOPPER["mass_body"] * 0.5 * h]) + rng.normal(0.0, 0.5, (B, 2))
    return np.ascontiguousarray(q1), np.ascontiguousarray(q2), np.ascontiguousarray(u)


def acrobot_batch(B, h=0.05, seed=0):
    """Elbow angle spread over the joint-limit range ±π/2 (some samples start on the limit)."""
    rng = np.random.default_rng(seed)
    q2 = np.stack([rng.uniform(-np.pi, np.pi, B), rng.uniform(-0.5 * np.pi, 0.5 * np.pi, B)], axis=1)
    q2[0::4, 1] = 0.5 * np.pi - 1e-3 * rng.uniform(0, 1, len(q2[0::4]))
    v = rng.normal(0.0, 1.0, (B, 2))
    q1 = q2 - h * v
    u = rng.normal(0.0, 0.5, (B, 1))      # control impulses; |u| ≳ 3 makes Newton wander for tens of iterations (chaotic, not comparable)
    return np.ascontiguousarray(q1), np.ascontiguousarray(q2), np.ascontiguousarray(u)


def cartpole_batch(B, h=0.05, seed=0):
    """Pendulum anywhere, a third of the samples at rest (stick regime of the joint friction)."""
    rng = np.random.default_rng(seed)
    q2 = np.stack([rng.uniform(-1.0, 1.0, B), rng.uniform(-np.pi, np.pi, B)], axis=1)
    v = rng.normal(0.0, 1.0, (B, 2))
    v[0::3] = 0.0
    q1 = q2 - h * v
    u = rng.normal(0.0, 1.0, (B, 1))
    u[0::3] *= 0.05
    return np.ascontiguousarray(q1), np.ascontiguousarray(q2), np.ascontiguousarray(u)


def planar_push_batch(B, h=0.1, seed=0):
    """Pusher near the −x face of the block (examples/planar_push.jl:46-47), half touching, half a few cm away."""
    rng = np.random.default_rng(seed)
    r_dim = 0.1
    pose = np.stack([rng.uniform(-0.2, 0.2, B), rng.uniform(-0.2, 0.2, B), rng.uniform(-0.5, 0.5, B)], axis=1)
    gap = rng.uniform(0.0, 0.05, B)
    gap[0::2] = 1.0e-8
    off = rng.uniform(-0.06, 0.06, B)
    c, s = np.cos(pose[:, 2]), np.sin(pose[:, 2])
    lx, ly = -r_dim - gap, off
    px = pose[:, 0] + c * lx - s * ly
    py = pose[:, 1] + s * lx + c * ly
    q2 = np.concatenate([pose, px[:, None], py[:, None]], axis=1)
    v = np.zeros((B, 5))
    v[:, 3:] = rng.normal(0.0, 0.05, (B, 2))
    v[1::2, :3] = rng.normal(0.0, 0.05, (len(v[1::2]), 3))
    q1 = q2 - h * v
    u = np.stack([rng.uniform(0.0, 1.0, B) * c, rng.uniform(0.0, 1.0, B) * s], axis=1) + rng.normal(0.0, 0.1, (B, 2))
    return np.ascontiguousarray(q1), np.ascontiguousarray(q2), np.ascontiguousarray(u)


def rocket_batch(B, seed=0, u_max=12.5):
    """States around the belly-flop start (examples/rocket.jl:44-55); thrust commands partly outside the SOC/u_max limits."""
    rng = np.random.default_rng(seed)
    x = np.zeros((B, 12))
    x[:, 0:3] = np.array([2.5, 2.5, 10.0]) + rng.normal(0.0, 1.0, (B, 3))
    x[:, 3:6] = rng.uniform(-0.4, 0.4, (B, 3))
    x[:, 6:9] = np.array([0.0, 0.0, -1.0]) + rng.normal(0.0, 1.0, (B, 3))
    x[:, 9:12] = rng.normal(0.0, 0.3, (B, 3))
    u = np.stack([rng.normal(0.0, 3.0, B), rng.normal(0.0, 3.0, B), rng.uniform(-2.0, 1.5 * u_max, B)], axis=1)
    return np.ascontiguousarray(x), np.ascontiguousarray(u)


def hopper_rollout_inputs(R, T=21, h=0.05, seed=0):
    """Forward-pass workload around the reference's hopper initial rollout (examples/hopper.jl:178,270-272): x1 = [q; q] with the
    foot on the ground, stand controls ū, plus a seeded feedback policy (K, k) and R step sizes α = 1, ½, ¼, … ≥ 1e-5
    (the Armijo candidates of examples/hopper.jl:276-278)."""
    rng = np.random.default_rng(seed)
    q = np.array([0.0, 0.5 + HOPPER["foot_radius"], 0.0, 0.5])
    x1 = np.concatenate([q, q])
    ubar = np.tile(np.array([0.0, HOPPER["gravity"] * HOPPER["mass_body"] * 0.5 * h]), (T - 1, 1)) + rng.normal(0.0, 0.05, (T - 1, 2))
    k = rng.normal(0.0, 0.3, (T - 1, 2))
    K = rng.normal(0.0, 0.2, (T - 1, 2, 8))
    alpha = np.maximum(0.5 ** np.arange(R), 1.0e-5)
    return x1, ubar, K, k, alpha


def planar_push_rollout_inputs(R, T=26, h=0.1, seed=0):
    """BASELINE.json configs[2]: planar push `rotate`, T = 26, R rollouts.  x1 and ū follow examples/planar_push.jl:46-47,112;
    every rollout gets its own seeded perturbation of the nominal controls."""
    rng = np.random.default_rng(seed)
    r_dim = 0.1
    q = np.array([0.0, 0.0, 0.0, -r_dim - 1.0e-8, -0.01])
    x1 = np.tile(np.concatenate([q, q]), (R, 1))
    nom = np.array([[1.0, 0.0] if t < 4 else [0.5, 0.0] if t < 9 else [0.0, 0.0] for t in range(T - 1)])
    ubar = nom[None] + rng.normal(0.0, 0.05, (R, T - 1, 2))
    return np.ascontiguousarray(x1), np.ascontiguousarray(ubar)


def quadratic_cost_expansion(X, U, x_goal, q_diag, r_diag, qT_diag, seed=None):
    """Cost expansion of the tracking objective the reference examples use (quadratic in x − x_goal and u, e.g.
    examples/cartpole.jl:50-66): lx, lu, lxx, luu (and a small seeded cross term lux when seed is given) along NT trajectories.
    X: [NT, T, n], U: [NT, T-1, m]."""
    NT, T, n = X.shape; m = U.shape[2]
    Q = np.diag(np.broadcast_to(q_diag, (n,)).astype(np.float64)); QT = np.diag(np.broadcast_to(qT_diag, (n,)).astype(np.float64))
    Rm = np.diag(np.broadcast_to(r_diag, (m,)).astype(np.float64))
    lxx = np.tile(Q, (NT, T, 1, 1)); lxx[:, -1] = QT
    lx = np.einsum("atij,atj->ati", lxx, X - np.asarray(x_goal)[None, None, :])
    luu = np.tile(Rm, (NT, T - 1, 1, 1))
    lu = np.einsum("atij,atj->ati", luu, U)
    lux = None
    if seed is not None:
        lux = 1.0e-2 * np.random.default_rng(seed).normal(size=(NT, T - 1, m, n))
    return lx, lu, lxx, luu, lux


def bundle_perturbations(ncol, N=64, eps=1.0e-4, seed=0):
    """One-hot perturbations η_i = ε·randn()·e_j (src/gradient_bundle.jl:49-54); the first ncol samples cover every coordinate
    once so the least-squares Hessian is never singular (the reference leaves that to chance)."""
    rng = np.random.default_rng(seed)
    eta = np.zeros((N, ncol))
    for i in range(N):
        j = i if i < ncol else int(rng.integers(0, ncol))
        eta[i, j] = eps * rng.normal()
    return eta
