"""mcmc_clv_model_b200 -- B200-native sampler for the Abe (2009/2015) hierarchical Pareto/NBD models.

Host-side mirror of the reference's entry points over a C-ABI CUDA library (libclv_b200.so, sm_100a).
Importing the package does not load the library; the first call does, and raises if it (or a CUDA
device) is missing -- there is no CPU fallback.
"""
from .api import (draw_future_transactions, draw_future_transactions_rfm_m, mcmc_draw_parameters,
                  mcmc_draw_parameters_rfm_m)
from .sampler import Sampler, default_hyper

__all__ = ["mcmc_draw_parameters", "mcmc_draw_parameters_rfm_m", "draw_future_transactions",
           "draw_future_transactions_rfm_m", "Sampler", "default_hyper"]
