#!/usr/bin/env python
"""How far do q3 and the IFT sensitivities move under each reading of the solver choices the reference tree does not pin?

RoboDojo.jl (the interior-point solver) is not in the reference tree; oracle/ip.hpp fixes four choices from recollection (its
header lists them).  This script re-runs the oracle with each alternative reading (oracle.set_variant) on the benchmark batches
and prints max / percentile |Δq3| and |Δ∂q3| against the default reading — the honest error bar on "parity" until
tests/golden/reference/*.csv exist (julia/dump_reference_golden.jl).  TEST INFRASTRUCTURE: uses oracle/ only.
usage: python tools/unpinned_sensitivity.py [hopper_batch=4096] > profiles/r02_unpinned_choice_sensitivity.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = {"tau = 1 - min(eps_min, vio^2)": dict(tau_rule=1), "tau = 0.99 fixed": dict(tau_rule=2), "reg = kappa_vio*gamma_reg applied": dict(apply_reg=1),
            "mu per cone dimension": dict(mu_mode=1), "SOC step tau <= 0.99": dict(soc_tau_cap=1)}


def study(name, B, seed=0):
    """rows: (variant, converged eval, converged grad, mean iterations, same-iteration-count fraction, |Δq3| median/p99/max, |Δgrad| median/p99/max)"""
    from oracle import oracle as O
    from common import CONFIGS
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u = gen(B, h=h, seed=seed)
    O.set_variant()
    e0 = O.step_batch(name, q1, q2, u, h, ke, False, fric=fric, diagnostics=False)
    g0 = O.step_batch(name, q1, q2, u, h, kg, True, fric=fric, diagnostics=False)
    rows = [("default", (e0["status"] == 0).mean(), (g0["status"] == 0).mean(), e0["iters"].mean(), 1.0, (0, 0, 0), (0, 0, 0))]
    try:
        for vn, kw in VARIANTS.items():
            O.set_variant(**kw)
            e = O.step_batch(name, q1, q2, u, h, ke, False, fric=fric, diagnostics=False)
            g = O.step_batch(name, q1, q2, u, h, kg, True, fric=fric, diagnostics=False)
            ok = (e0["status"] == 0) & (e["status"] == 0); okg = (g0["status"] == 0) & (g["status"] == 0)
            dq = np.abs(e["q3"] - e0["q3"]).max(1)[ok]
            dg = np.maximum.reduce([np.abs(g[k] - g0[k]).reshape(B, -1).max(1) for k in ("dq1", "dq2", "du")])[okg]
            q = lambda a: (float(np.median(a)), float(np.quantile(a, 0.99)), float(a.max()))       # noqa: E731
            rows.append((vn, (e["status"] == 0).mean(), (g["status"] == 0).mean(), e["iters"].mean(), (e["iters"] == e0["iters"]).mean(), q(dq), q(dg)))
    finally:
        O.set_variant()
    return rows


def main():
    Bh = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    print(__doc__.split("usage:")[0])
    for name, B in (("hopper", Bh), ("cartpole_friction", Bh), ("acrobot_impact", Bh), ("planar_push", min(Bh, 1024))):
        print("%s, %d problems (benchmark batch, seed 0)" % (name, B))
        print("  %-36s %9s %9s %6s %9s | %-32s | %-32s" % ("reading", "conv eval", "conv grad", "iters", "same #it", "|dq3| median / p99 / max", "|d grad| median / p99 / max"))
        for vn, ce, cg, it, same, dq, dg in study(name, B):
            print("  %-36s %9.4f %9.4f %6.2f %9.3f | %9.2e %9.2e %9.2e | %9.2e %9.2e %9.2e" % ((vn, ce, cg, it, same) + tuple(dq) + tuple(dg)))
        print()


if __name__ == "__main__":
    main()
