"""Minimal stand-ins for the GPJax objects the CaGP hot path consumes.

GPJax (``gpjax>=0.14.0rc1``) is a third-party dependency of the reference that is not vendored
in /root/reference and not installable here.  The reference's model layer receives GPJax objects
(``Prior``, ``ConjugatePosterior``, ``Gaussian``, ``Dataset``, kernels with a ``compute_engine``,
mean functions; models/cagp.py:14-16, README.md:29-47).  This module provides objects with the
same names, constructor arguments and attributes, backed by torch tensors, so that code written
against the reference reads the same.  Only what the hot path touches is implemented.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Callable, Iterable, List, Optional

import torch

DEFAULT_DTYPE = torch.float64  # the reference runs with jax_enable_x64 (README.md:24-25)


# --------------------------------------------------------------------------------------------
# Parameters (gpjax.parameters.Real / PositiveReal, paramax.non_trainable / unwrap)
# --------------------------------------------------------------------------------------------
class Parameter:
    """A leaf of the model.  ``value`` is a torch tensor; ``trainable`` marks it for ``fit``;
    ``positive`` makes ``fit`` optimise it through a softplus reparametrisation."""

    positive = False

    def __init__(self, value, trainable: bool = True, dtype=None):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value, dtype=dtype or DEFAULT_DTYPE)
        self.value = value
        self.trainable = trainable

    def __repr__(self):
        return f"{type(self).__name__}({self.value!r}, trainable={self.trainable})"


class Real(Parameter):
    pass


class PositiveReal(Parameter):
    positive = True


class NonTrainable(Parameter):
    def __init__(self, value, dtype=None):
        inner = value.value if isinstance(value, Parameter) else value
        super().__init__(inner, trainable=False, dtype=dtype)


def non_trainable(value) -> NonTrainable:
    """paramax.non_trainable"""
    return NonTrainable(value)


def unwrap(x):
    """paramax.unwrap: return the raw tensor behind a (possibly wrapped) leaf."""
    return x.value if isinstance(x, Parameter) else x


def _as_param(v, cls):
    if isinstance(v, Parameter):
        return v
    return cls(v)


# --------------------------------------------------------------------------------------------
# Dataset
# --------------------------------------------------------------------------------------------
@dataclass
class Dataset:
    X: Optional[torch.Tensor] = None
    y: Optional[torch.Tensor] = None

    @property
    def n(self) -> int:
        return 0 if self.X is None else self.X.shape[0]


# --------------------------------------------------------------------------------------------
# Mean functions
# --------------------------------------------------------------------------------------------
class AbstractMeanFunction:
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class Constant(AbstractMeanFunction):
    def __init__(self, constant=0.0):
        self.constant = constant if isinstance(constant, Parameter) or torch.is_tensor(constant) else Real(constant)

    def __call__(self, x):
        c = unwrap(self.constant)
        return torch.ones((x.shape[0], 1), dtype=c.dtype, device=x.device) * c.to(x.device)


class Zero(Constant):
    def __init__(self):
        self.constant = NonTrainable(torch.zeros((), dtype=DEFAULT_DTYPE))

    def __call__(self, x):
        return torch.zeros((x.shape[0], 1), dtype=x.dtype, device=x.device)


# --------------------------------------------------------------------------------------------
# Kernels
# --------------------------------------------------------------------------------------------
class AbstractKernelComputation:
    """gpjax.kernels.computations.AbstractKernelComputation protocol."""

    def gram(self, kernel, x):
        raise NotImplementedError

    def cross_covariance(self, kernel, x1, x2):
        raise NotImplementedError


class AbstractKernel:
    """Stationary kernel with ``lengthscale`` (scalar or (d,)) and ``variance`` leaves."""

    name = "abstract"  # family key understood by the native kernels

    def __init__(self, lengthscale=1.0, variance=1.0, compute_engine: Optional[AbstractKernelComputation] = None,
                 n_dims: Optional[int] = None):
        self.lengthscale = _as_param(lengthscale, PositiveReal)
        self.variance = _as_param(variance, PositiveReal)
        if compute_engine is None:
            from .computations import LazyKernelComputation  # the only engine of this package

            compute_engine = LazyKernelComputation()
        self.compute_engine = compute_engine
        self.n_dims = n_dims

    # single pair evaluation, used for the dtype probe at operators/lazy_kernel.py:49
    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        ls = unwrap(self.lengthscale).to(x.device)
        var = unwrap(self.variance).to(x.device)
        sq = (((x - y) / ls) ** 2).sum()
        return var * self._profile(sq)

    def _profile(self, sq):
        raise NotImplementedError

    def gram(self, x):
        return self.compute_engine.gram(self, x)

    def cross_covariance(self, x1, x2):
        return self.compute_engine.cross_covariance(self, x1, x2)


class RBF(AbstractKernel):
    name = "rbf"

    def _profile(self, sq):
        return torch.exp(-0.5 * sq)


def _tau(sq):
    return torch.sqrt(torch.clamp(sq, min=1e-36))


class Matern12(AbstractKernel):
    name = "matern12"

    def _profile(self, sq):
        return torch.exp(-_tau(sq))


class Matern32(AbstractKernel):
    name = "matern32"

    def _profile(self, sq):
        r = math.sqrt(3.0) * _tau(sq)
        return (1.0 + r) * torch.exp(-r)


class Matern52(AbstractKernel):
    name = "matern52"

    def _profile(self, sq):
        t = _tau(sq)
        r = math.sqrt(5.0) * t
        return (1.0 + r + 5.0 / 3.0 * t * t) * torch.exp(-r)


# --------------------------------------------------------------------------------------------
# Likelihood, prior, posterior
# --------------------------------------------------------------------------------------------
class Gaussian:
    """gpjax.likelihoods.Gaussian (analytical integrator)."""

    def __init__(self, num_datapoints: int, obs_stddev=1.0):
        self.num_datapoints = num_datapoints
        self.obs_stddev = _as_param(obs_stddev, PositiveReal) if not torch.is_tensor(obs_stddev) else obs_stddev

    def expected_log_likelihood(self, y, mean, variance):
        """E_q[log N(y | f, sigma^2)] per point; y, mean, variance are (n, 1) (call site cagp.py:227-229)."""
        sigma = unwrap(self.obs_stddev).to(mean.device)
        sq = sigma**2
        y = y.to(mean.device)
        val = -0.5 * (math.log(2.0 * math.pi) + torch.log(sq) + ((y - mean) ** 2 + variance) / sq)
        return val.sum(dim=1)


class Prior:
    def __init__(self, kernel: AbstractKernel, mean_function: AbstractMeanFunction, jitter: float = 1e-6):
        self.kernel = kernel
        self.mean_function = mean_function
        self.jitter = jitter

    def __mul__(self, likelihood: Gaussian) -> "ConjugatePosterior":
        return ConjugatePosterior(prior=self, likelihood=likelihood)


class ConjugatePosterior:
    def __init__(self, prior: Prior, likelihood: Gaussian, jitter: float = 1e-6):
        self.prior = prior
        self.likelihood = likelihood
        self.jitter = jitter


# --------------------------------------------------------------------------------------------
# Parameter traversal and fit (gpx.fit; README.md:57-68)
# --------------------------------------------------------------------------------------------
def _leaf_slots(model) -> List[tuple]:
    """(owner, attribute) pairs of every model leaf, in a fixed order."""
    post = model.posterior
    slots = [(post.prior.kernel, "lengthscale"), (post.prior.kernel, "variance"), (post.likelihood, "obs_stddev")]
    if isinstance(post.prior.mean_function, Constant) and not isinstance(post.prior.mean_function, Zero):
        slots.append((post.prior.mean_function, "constant"))
    policy = model.policy
    while policy is not None:  # wrappers (OrthogonalizationPolicy) hold their leaves in ``base_policy``
        for attr in ("nz_values", "actions", "pseudo_inputs"):
            if hasattr(policy, attr):
                slots.append((policy, attr))
        inner_kernel = getattr(policy, "kernel", None)
        if inner_kernel is not None and inner_kernel is not post.prior.kernel:
            slots += [(inner_kernel, "lengthscale"), (inner_kernel, "variance")]
        policy = getattr(policy, "base_policy", None)
    return slots


def trainable_leaves(model) -> List[tuple]:
    out = []
    for owner, attr in _leaf_slots(model):
        leaf = getattr(owner, attr)
        if isinstance(leaf, Parameter):
            if leaf.trainable and leaf.value.is_floating_point():
                out.append((owner, attr, leaf))
        elif torch.is_tensor(leaf) and leaf.is_floating_point():
            out.append((owner, attr, leaf))
    return out


def value_and_grad(objective: Callable, model, train_data: Dataset):
    """``jax.value_and_grad(objective)(model, train_data)``: returns (value, {leaf name: grad})."""
    leaves = trainable_leaves(model)
    tensors = []
    for owner, attr, leaf in leaves:
        t = unwrap(leaf).detach().clone().requires_grad_(True)
        tensors.append(t)
        if isinstance(leaf, Parameter):
            leaf.value = t
        else:
            setattr(owner, attr, t)
    val = objective(model, train_data)
    grads = torch.autograd.grad(val, tensors, allow_unused=True)
    names = [f"{type(o).__name__}.{a}" for o, a, _ in leaves]
    out = {}
    for name, t, g in zip(names, tensors, grads):
        out[name] = torch.zeros_like(t) if g is None else g
    for (owner, attr, leaf), t in zip(leaves, tensors):
        if isinstance(leaf, Parameter):
            leaf.value = t.detach()
        else:
            setattr(owner, attr, t.detach())
    return val.detach(), out


def _softplus_inv(x):
    return x + torch.log(-torch.expm1(-x))


def fit(*, model, objective: Callable, train_data: Dataset, optim="adamw", num_iters: int = 100,
        learning_rate: float = 0.01, key=None, verbose: bool = False, cuda_graph: bool = False):
    """Gradient-based optimisation of ``objective(model, train_data)`` over the trainable leaves.

    Mirrors ``gpx.fit(model=, objective=, train_data=, optim=, num_iters=, key=)``: positive
    parameters are optimised in an unconstrained (softplus) space, as GPJax does.  ``optim`` is
    "adamw", "adam", "sgd" or a callable ``params -> torch.optim.Optimizer``.  Returns
    ``(model, history)`` with the per-iteration objective values.

    ``cuda_graph=True`` (single GPU, built-in optimisers): after three eager iterations the whole step - forward
    kernels, k x k algebra, reverse mode, optimiser update - is captured once in a CUDA graph and replayed.  A step
    of a small problem (README shape: n = 1000, k = 10) is ~100 launches of microsecond kernels, i.e. bound by
    launch latency and Python; the replay removes both (what ``jax.jit`` does for the reference's training loop).
    """
    leaves = trainable_leaves(model)
    raw = []
    for owner, attr, leaf in leaves:
        v = unwrap(leaf).detach().clone()
        if cuda_graph and train_data.X is not None:
            v = v.to(train_data.X.device)  # every leaf the captured step updates must live on the GPU
        if isinstance(leaf, Parameter) and leaf.positive:
            v = _softplus_inv(v)
        raw.append(v.requires_grad_(True))
    if cuda_graph and (callable(optim) or not raw or not all(r.is_cuda for r in raw)):
        raise ValueError("cuda_graph=True needs CUDA training data and one of the built-in optimisers")
    if callable(optim):
        opt = optim(raw)
    elif optim == "adamw":
        opt = torch.optim.AdamW(raw, lr=learning_rate, weight_decay=1e-4, capturable=cuda_graph)
    elif optim == "adam":
        opt = torch.optim.Adam(raw, lr=learning_rate, capturable=cuda_graph)
    elif optim == "sgd":
        opt = torch.optim.SGD(raw, lr=learning_rate)
    else:
        raise ValueError(f"unknown optimiser {optim!r}")
    if cuda_graph and optim in ("adam", "adamw"):
        # capturable Adam would create its step counter (hence the bias corrections 1 - beta^t) in the process-wide
        # DEFAULT dtype; float32 there perturbs a float64 trajectory at the 1e-6 level.  Create the state here, in the
        # leaves' own dtype, instead of changing the global default around the fit.
        for r in raw:
            opt.state[r] = {"step": torch.zeros((), dtype=r.dtype, device=r.device),
                            "exp_avg": torch.zeros_like(r, memory_format=torch.preserve_format),
                            "exp_avg_sq": torch.zeros_like(r, memory_format=torch.preserve_format)}

    def install(detach: bool):
        for (owner, attr, leaf), r in zip(leaves, raw):
            v = torch.nn.functional.softplus(r) if isinstance(leaf, Parameter) and leaf.positive else r
            if detach:
                v = v.detach().clone()
            if isinstance(leaf, Parameter):
                leaf.value = v
            else:
                setattr(owner, attr, v)

    def step():
        install(detach=False)
        loss = objective(model, train_data)
        loss.backward()
        opt.step()
        return loss

    history = []
    n_eager = min(num_iters, 3) if cuda_graph else num_iters
    if cuda_graph:  # warm-up on a side stream, as torch.cuda.graph requires
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)  # the leaves were created on the default stream; the mismatch is intended here
        side = torch.cuda.Stream(device=raw[0].device)
        side.wait_stream(torch.cuda.current_stream(raw[0].device))
        ctx = torch.cuda.stream(side)
    else:
        import contextlib

        ctx = contextlib.nullcontext()
    with ctx:
        for it in range(n_eager):
            opt.zero_grad(set_to_none=True)
            loss = step()
            history.append(loss.detach().clone() if cuda_graph else loss.detach())
            if verbose and (it % 10 == 0 or it == num_iters - 1):
                print(f"iter {it}: objective {float(loss):.6f}")
    if cuda_graph and num_iters > n_eager:
        torch.cuda.current_stream(raw[0].device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = step()
        for it in range(n_eager, num_iters):
            graph.replay()
            history.append(static_loss.detach().clone())
            if verbose and (it % 10 == 0 or it == num_iters - 1):
                print(f"iter {it}: objective {float(static_loss):.6f}")
    install(detach=True)
    return model, torch.stack(history) if history else torch.zeros(0)
