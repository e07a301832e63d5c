"""Drop-in for the reference's `src/models/utils/elog2cbs2param.py`: `elog2cbs` executed on the device."""
from mcmc_clv_model_b200.cbs import elog2cbs

__all__ = ["elog2cbs"]
