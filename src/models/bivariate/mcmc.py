"""Drop-in for the reference's `src/models/bivariate/mcmc.py` (same import path, same `__all__`):
the Abe (2009) hierarchical-Bayes Pareto/NBD sampler and forecast, executed by the B200 CUDA library.

    from src.models.bivariate.mcmc import mcmc_draw_parameters, draw_future_transactions

works exactly as in the reference's run_mcmc_*.py / analysis_*.py (bi:437, bi:506).  `draw_z` and
`draw_tau` are exported by the reference too (bi:45-46); here they are single-block device calls that
take the uniforms/exponentials explicitly drawn from the given NumPy generator in the reference's order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from mcmc_clv_model_b200.api import draw_future_transactions, mcmc_draw_parameters
from mcmc_clv_model_b200.blocks import draw_tau, draw_z
from mcmc_clv_model_b200.synthetic import elog2cbs, generate_pareto_abe

__all__ = [
    "CustomerCBS",
    "elog2cbs",
    "generate_pareto_abe",
    "draw_z",
    "draw_tau",
    "mcmc_draw_parameters",
    "draw_future_transactions",
]


@dataclass
class CustomerCBS:
    """Sufficient statistics for one customer in the calibration window (bi:55-69)."""

    x: int
    t_x: float
    T_cal: float

    @property
    def frequency(self) -> int:
        return self.x

    @property
    def recency(self) -> float:
        return self.t_x
