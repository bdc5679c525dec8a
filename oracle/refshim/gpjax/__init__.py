"""Stand-in for `gpjax` (>= 0.14): Dataset, parameters, constant mean, stationary kernels with pluggable compute
engines, Gaussian likelihood, Prior / ConjugatePosterior.  Formulas restated from GPJax's published definitions (the
same ones tests/test_oracle_third_party_anchors.py checks against scikit-learn)."""
from . import dataset, distributions, gps, kernels, likelihoods, mean_functions, parameters  # noqa: F401
from .dataset import Dataset  # noqa: F401
from .gps import ConjugatePosterior, Prior  # noqa: F401
