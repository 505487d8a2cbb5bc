"""optimization_dynamics_b200 — B200-native contact-implicit step + IFT gradients behind the reference's f / fx / fu API.

Exports mirror reference src/OptimizationDynamics.jl:28-34,75-88."""
from .dynamics import (ImplicitDynamics, Simulator, Model, f, fx, fu, state_to_configuration, acrobot_impact, acrobot_nominal, cartpole_friction,
                       cartpole_frictionless, planarpush, hopper, rocket)
from .gradient_bundle import GradientBundle, gradient, gradient_batch, fx_gb, fu_gb
from .rocket import (RocketInfo, f_rocket, fx_rocket, fu_rocket, soc_projection, soc_projection_gradient, f_rocket_proj, fx_rocket_proj,
                     fu_rocket_proj)
from .rollout import rollout, rollout_batch
from .riccati import backward_pass_batch
from . import workloads
from . import robodojo

__all__ = ["ImplicitDynamics", "Model", "f", "fx", "fu", "state_to_configuration", "GradientBundle", "gradient", "gradient_batch", "fx_gb",
           "fu_gb", "RocketInfo", "f_rocket", "fx_rocket", "fu_rocket", "soc_projection", "soc_projection_gradient", "f_rocket_proj",
           "fx_rocket_proj", "fu_rocket_proj", "acrobot_impact", "acrobot_nominal", "cartpole_friction", "cartpole_frictionless",
           "planarpush", "hopper", "rocket", "rollout", "rollout_batch", "backward_pass_batch", "workloads", "robodojo", "Simulator"]
