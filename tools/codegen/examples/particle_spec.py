"""Example model specification for tools/codegen/gen_models.py --spec: a point mass in the x–z plane with a ground contact
(z ≥ 0) and Coulomb friction along x — the smallest model that exercises every block of the contact-implicit step.

    python tools/codegen/gen_models.py --spec tools/codegen/examples/particle_spec.py --out /tmp/gen

A specification is a Python module with a function `model()` returning a dict:

    name                 identifier; the generated code lives in namespace od::gen_<name>, the traits struct is od::<Name>Model
    NQ, NU               configuration and control dimensions
    NC                   orthant pairs (γ_i, s_i): impact / distance constraints   s = ϕ(q) ≥ 0, γ ≥ 0
    NP, NB, cone_dims    friction cones: NP cones (ψ_k; b_k) with cone_dims[k] tangential components each, NB = Σ cone_dims
    NTH, th              data vector θ = [q0 (NQ), q1 (NQ), u (NU), friction parameters (NTH − 2NQ − NU − 1), h]
    q, gam, b            sympy symbols of the unknown configuration q2, the normal impulses γ and the friction impulses b
    d                    NQ dynamics rows  d(q, γ, b; θ) = 0      (variational integrator + control + Jᵀ[γ; b])
    phi                  NC signed distances ϕ(q)
    psit                 NP friction-cone radii  ψ̂(γ; θ)  (μ γ)
    vT                   NB tangential velocities vT(q; θ)

Everything else — exact Jacobians, common-subexpression elimination, the sin/cos and root tables, the C++ traits struct that binds
the code to the solver templates — is produced by the generator.  (Successor of the reference's per-model `codegen.jl` +
`deps/build.jl`; to ship a model in liboptdyn_b200.so, add its header to csrc/models.cuh and a case to the dispatch in
csrc/optdyn_b200.cu.)
"""
import sympy as sp

from gen_models import vec


def model():
    NQ, NU, NC, NP, NB = 2, 2, 1, 1, 1
    q, gam, b = vec("q", NQ), vec("g", NC), vec("b", NB)
    th = vec("t", 2 * NQ + NU + 1 + 1)
    q0, q1, u, mu, h = th[0:2], th[2:4], th[4:6], th[6], th[7]
    mass, grav = 1.5, 9.81
    vm1 = [(q1[i] - q0[i]) / h for i in range(2)]
    vm2 = [(q[i] - q1[i]) / h for i in range(2)]
    force = [u[0], u[1] - mass * grav]
    contact = [b[0], gam[0]]                                   # Jᵀ[γ; b]: friction along x, normal along z
    d = [mass * (vm1[i] - vm2[i]) + h * force[i] + contact[i] for i in range(2)]
    return dict(name="particle", NQ=NQ, NU=NU, NC=NC, NP=NP, NB=NB, cone_dims=[1], NTH=len(th), q=q, gam=gam, b=b, th=th,
                d=d, phi=[q[1]], psit=[mu * gam[0]], vT=[vm2[0]])
