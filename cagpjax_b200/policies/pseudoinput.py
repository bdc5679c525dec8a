"""Pseudo-input policy (reference: policies/pseudoinput.py:14-74).

Actions are the cross-covariance S = K(X, Z) between the training inputs and (trainable) pseudo-inputs in the
same input space.  The (n, M) matrix is materialised by the native cross-covariance kernel
(``cagp_cross_cov_fwd``; reverse mode w.r.t. Z, the lengthscale and the variance through ``cagp_cross_cov_bwd``)
and returned as a dense operator, which ``ComputationAwareGP`` then feeds to the dense-action kernels."""

from __future__ import annotations

import torch

from .. import _fused
from ..gpx import AbstractKernel, Dataset, NonTrainable, unwrap
from ..ops import Dense
from .base import AbstractBatchLinearSolverPolicy


class PseudoInputPolicy(AbstractBatchLinearSolverPolicy):
    """Args:
        pseudo_inputs: (M, D) tensor or parameter (wrap in ``gpx.Real`` to train them).
        train_inputs_or_dataset: the training inputs (N, D), or a dataset holding them - the same inputs, in the
            same order, as the data the model is conditioned on.
        kernel: kernel of the GP prior.

    Many pseudo-inputs make S poorly conditioned; wrap the policy in ``OrthogonalizationPolicy`` then."""

    def __init__(self, pseudo_inputs, train_inputs_or_dataset, kernel: AbstractKernel):
        self.pseudo_inputs = pseudo_inputs
        if isinstance(train_inputs_or_dataset, Dataset):
            if train_inputs_or_dataset.X is None:
                raise ValueError("Dataset must contain training inputs.")
            train_inputs = train_inputs_or_dataset.X
        else:
            train_inputs = train_inputs_or_dataset
        self.train_inputs = NonTrainable(torch.atleast_2d(torch.as_tensor(train_inputs)))
        self.kernel = kernel
        self.n_actions = unwrap(self.pseudo_inputs).shape[0]
        if unwrap(self.train_inputs).shape[1:] != unwrap(self.pseudo_inputs).shape[1:]:
            raise ValueError("Training inputs and pseudo-inputs must have the same trailing dimensions.")
        self._check_init()

    def to_actions(self, A, *, key=None) -> Dense:
        X = unwrap(self.train_inputs)
        Z = unwrap(self.pseudo_inputs)
        if A is not None and hasattr(A, "device") and X.device != A.device:
            X = X.to(A.device)
        Z = Z.to(device=X.device, dtype=X.dtype)
        ls = unwrap(self.kernel.lengthscale).to(device=X.device, dtype=X.dtype)
        var = unwrap(self.kernel.variance).to(device=X.device, dtype=X.dtype)
        return Dense(var * _fused.cross_covariance(self.kernel.name, X, Z, ls))
