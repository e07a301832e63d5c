"""Drop-in for the reference's `src/models/trivariate/mcmc.py`: the 3-parameter (lambda, mu, eta)
"RFM-M" sampler (tri:580) and its forecast with log-normal spend (tri:660), executed by the B200
CUDA library.

    from src.models.trivariate.mcmc import mcmc_draw_parameters_rfm_m, draw_future_transactions
"""
from __future__ import annotations

from mcmc_clv_model_b200.api import draw_future_transactions_rfm_m as draw_future_transactions
from mcmc_clv_model_b200.api import mcmc_draw_parameters_rfm_m
from mcmc_clv_model_b200.blocks import draw_tau, draw_z

__all__ = ["draw_z", "draw_tau", "mcmc_draw_parameters_rfm_m", "draw_future_transactions"]
