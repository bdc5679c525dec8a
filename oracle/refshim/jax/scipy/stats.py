"""Stand-in for `jax.scipy.stats` (scipy's own distributions)."""
from scipy.stats import multivariate_normal, norm  # noqa: F401
