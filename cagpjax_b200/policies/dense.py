"""Dense action-matrix policy.

The reference has no policy class for a user-supplied dense action matrix (its dense producers are
``LanczosPolicy`` / ``PseudoInputPolicy``, which return ``cola.ops.Dense``); BASELINE config 5 ("dense
random-Gaussian action matrix") is expressed through the ``AbstractBatchLinearSolverPolicy`` interface
(reference: policies/base.py:18-43) by this minimal policy returning a ``Dense`` operator."""

from __future__ import annotations

import torch

from ..gpx import unwrap
from ..ops import Dense
from .base import AbstractBatchLinearSolverPolicy


class DensePolicy(AbstractBatchLinearSolverPolicy):
    """Actions given by a dense, learnable (n, n_actions) matrix.

    ``fused_statistics``: the matrix is expected to be well-conditioned (random Gaussian, orthogonal, ...), which
    lets ``ComputationAwareGP`` evaluate the ELBO from the k-sized statistics of one fused pass (see
    ``models.cagp._fused_dense_statistics_ok``); set it to False for ill-conditioned action matrices."""

    fused_statistics = True

    def __init__(self, actions):
        self.actions = actions
        self.n_actions = unwrap(actions).shape[1]
        self._check_init()

    @classmethod
    def from_random(cls, key, num_datapoints: int, n_actions: int, *, dtype=None, device=None) -> "DensePolicy":
        """i.i.d. N(0, 1/n) entries."""
        if num_datapoints < 1:
            raise ValueError("num_datapoints must be at least 1")
        gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
        S = torch.randn((num_datapoints, n_actions), generator=gen, dtype=dtype or torch.float64) / num_datapoints**0.5
        return cls(S.to(device) if device is not None else S)

    def to_actions(self, A, *, key=None) -> Dense:
        S = unwrap(self.actions)
        if A is not None and hasattr(A, "device") and S.device != A.device:
            S = S.to(A.device)
        return Dense(S)
