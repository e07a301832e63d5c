"""Variate sources for the oracle (TEST INFRASTRUCTURE ONLY).

The reference threads one ``numpy.random.Generator`` through every block
(bi:193,203,233,268; Appendix B of SURVEY.md gives the call order).  The oracle
asks one of these sources instead:

* ``NumpyOrderStreams`` – NumPy PCG64 consumed in exactly the reference's call
  order and through the same third-party entry points (``rng.random``,
  ``rng.exponential``, ``rng.standard_t``, ``scipy.stats.invwishart.rvs``,
  ``rng.multivariate_normal``) ⇒ the oracle reproduces the reference's chains.
* ``ReplayStreams`` – caller-supplied full-length arrays (injected-stream
  parity; also what the GPU's injected mode consumes).
* ``PhiloxStreams`` – the device's counter-based contract (``philox_np``).
"""
from __future__ import annotations

import numpy as np

from . import abe_oracle as ao
from . import philox_np as px


class _Base:
    def begin_sweep(self, step):  # 1-based sweep number, as in bi:383
        self.step = step

    def inv_wishart(self, nu_n, S_n):
        v = self._level2_variates(S_n.shape[0], None, nu_n)
        return ao.inv_wishart_bartlett(S_n, v["iw_norm"], v["iw_chi2"])

    def beta(self, B_hat, Sigma, V, compat):
        K, D = B_hat.shape
        v = self._level2_variates(D, K, None)
        return ao.assemble_beta(B_hat, ao.beta_noise_cholesky(Sigma, V, v["beta_norm"]), compat)


class NumpyOrderStreams(_Base):
    def __init__(self, rng: np.random.Generator):
        self.rng = rng

    def u_z(self, N):
        return self.rng.random(N)

    def e_tau(self, alive):
        out = np.zeros(alive.shape)
        n = int(alive.sum())
        if n:
            out[alive] = self.rng.standard_exponential(n)
        return out

    def u_tau(self, churn):
        out = np.zeros(churn.shape)
        n = int(churn.sum())
        if n:
            out[churn] = self.rng.random(n)
        return out

    def t3(self, N):
        return self.rng.standard_t(df=3, size=N)

    def u_acc(self, N):
        return self.rng.random(N)

    def n_eta(self, N):
        return self.rng.standard_normal(N)

    def inv_wishart(self, nu_n, S_n):
        from scipy.stats import invwishart
        return invwishart.rvs(df=nu_n, scale=S_n, random_state=self.rng)

    def beta(self, B_hat, Sigma, V, compat):
        assert compat == "reference"
        return self.rng.multivariate_normal(B_hat.ravel(), np.kron(Sigma, V)).reshape(B_hat.shape)


class ReplayStreams(_Base):
    """arrays: dict with a leading sweep axis T (sweep ``step`` uses row step-1):
    u_z, e_tau, u_tau (T,N); t3_l, t3_m, u_acc (T,S,N); n_eta (T,N) [D=3];
    iw_norm (T,D(D-1)/2), iw_chi2 (T,D), beta_norm (T,D*K)."""

    def __init__(self, arrays):
        self.a = arrays

    def begin_sweep(self, step):
        self.step = step
        self._t3_calls = 0
        self._acc_calls = 0

    def _row(self, name):
        return np.asarray(self.a[name][self.step - 1])

    def u_z(self, N):
        return self._row("u_z")

    def e_tau(self, alive):
        return self._row("e_tau")

    def u_tau(self, churn):
        return self._row("u_tau")

    def t3(self, N):
        s, which = divmod(self._t3_calls, 2)
        self._t3_calls += 1
        return self._row("t3_l" if which == 0 else "t3_m")[s]

    def u_acc(self, N):
        s = self._acc_calls
        self._acc_calls += 1
        return self._row("u_acc")[s]

    def n_eta(self, N):
        return self._row("n_eta")

    def _level2_variates(self, D, K, nu_n):
        return dict(iw_norm=self._row("iw_norm"), iw_chi2=self._row("iw_chi2"),
                    beta_norm=self._row("beta_norm"))


class PhiloxStreams(_Base):
    """Strict-f64 device contract; ``gids`` are global customer ids."""

    def __init__(self, seed, chain, gids, n_mh_steps, D, K, level1_variates=None):
        """level1_variates: optional callable(sweep) -> dict(t3_l, t3_m, u_acc), each (S, N), replacing the strict-f64
        Metropolis variates -- the hook through which a test replays the variates the device's FAST transforms
        produced (clv_debug_variates) so that the FAST kernel's whole trajectory can be checked."""
        self.seed, self.chain = seed, chain
        self.gids = np.asarray(gids)
        self.S, self.D, self.K = n_mh_steps, D, K
        self._cache_step = None
        self._level1_variates = level1_variates

    def begin_sweep(self, step):
        self.step = step
        self._t3_calls = 0
        self._acc_calls = 0
        self._v = px.sampler_variates(self.seed, self.chain, self.gids, step, self.S,
                                      with_eta=(self.D == 3))
        if self._level1_variates is not None:
            self._v.update(self._level1_variates(step))
        self._l2 = None

    def u_z(self, N):
        return self._v["u_z"]

    def e_tau(self, alive):
        return self._v["e_tau"]

    def u_tau(self, churn):
        return self._v["u_tau"]

    def t3(self, N):
        s, which = divmod(self._t3_calls, 2)
        self._t3_calls += 1
        return self._v["t3_l" if which == 0 else "t3_m"][s]

    def u_acc(self, N):
        s = self._acc_calls
        self._acc_calls += 1
        return self._v["u_acc"][s]

    def n_eta(self, N):
        return self._v["n_eta"]

    def inv_wishart(self, nu_n, S_n):
        self._l2 = px.level2_variates(self.seed, self.chain, self.step, self.D, self.K, nu_n)
        return ao.inv_wishart_bartlett(S_n, self._l2["iw_norm"], self._l2["iw_chi2"])

    def _level2_variates(self, D, K, nu_n):
        return self._l2


def random_replay_arrays(rng, T, N, S, D, K, nu_n):
    """Convenience: a full set of injected variates drawn from ``rng``."""
    a = dict(
        u_z=rng.random((T, N)), e_tau=rng.standard_exponential((T, N)), u_tau=rng.random((T, N)),
        t3_l=rng.standard_t(3, (T, S, N)), t3_m=rng.standard_t(3, (T, S, N)),
        u_acc=rng.random((T, S, N)),
        iw_norm=rng.standard_normal((T, D * (D - 1) // 2)),
        iw_chi2=np.stack([rng.chisquare(nu_n - D + 1 + i, T) for i in range(D)], axis=1),
        beta_norm=rng.standard_normal((T, D * K)),
    )
    if D == 3:
        a["n_eta"] = rng.standard_normal((T, N))
    return a
