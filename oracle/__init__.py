"""CPU oracle for the Abe (2009/2015) sampler hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product (``mcmc_clv_model_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the reference's
own NumPy code (imported read-only from ``/root/reference`` in the build
container) both with its real PCG64 streams and with injected streams, and the
committed fixtures under ``tests/golden/`` hold its outputs;
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
