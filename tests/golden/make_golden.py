"""Generates tests/golden/*.npz from the CPU oracle (oracle/) on seeded inputs.

The reference's own implementation cannot be executed in this environment (no Julia runtime; RoboDojo.jl is an un-vendored
dependency), and the reference ships no golden vectors — so these fixtures are ORACLE outputs ("parity unpinned").  They exist
so that (a) the oracle cannot drift silently and (b) the GPU tests have fixed vectors that do not depend on building the oracle
on the GPU box.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O          # noqa: E402
from optimization_dynamics_b200 import workloads as W   # noqa: E402
from common import CONFIGS              # noqa: E402

B = 96
REF_IN = os.path.join(HERE, "reference_inputs")       # the same inputs as text, for julia/dump_reference_golden.jl (17 significant digits: exact)
os.makedirs(REF_IN, exist_ok=True)
for name, (gen, h, ke, kg, fric, _) in CONFIGS.items():
    q1, q2, u = gen(B, h=h, seed=2024)
    np.savetxt(os.path.join(REF_IN, name + ".csv"), np.concatenate([q1, q2, u], axis=1), delimiter=",", fmt="%.17g")
    e = O.step_batch(name, q1, q2, u, h, ke, False, fric=fric)
    g = O.step_batch(name, q1, q2, u, h, kg, True, fric=fric)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), q1=q1, q2=q2, u=u, q3=e["q3"], dq1=g["dq1"], dq2=g["dq2"], du=g["du"],
                        status_eval=e["status"], status_grad=g["status"], iters_eval=e["iters"], iters_grad=g["iters"],
                        margin=np.minimum(e["margin"], g["margin"]), ift_spread=g["ift_spread"],
                        q_uncertainty=np.maximum(e["q_uncertainty"], g["q_uncertainty"]))
x, u = W.rocket_batch(B, seed=2024)
np.savetxt(os.path.join(REF_IN, "rocket.csv"), np.concatenate([x, u], axis=1), delimiter=",", fmt="%.17g")
for proj, name in ((False, "rocket"), (True, "rocket_proj")):
    r = O.rocket_batch(x, u, 0.05, 12.5, proj, True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, u=u, y=r["y"], dx=r["dx"], du=r["du"], uproj=r["uproj"], status=r["status"],
                        margin=r["margin"])
print("golden vectors written to", HERE)
