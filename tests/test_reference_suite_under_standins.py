"""The reference's OWN test suite, run against its unmodified source under the dependency stand-ins (CPU, this
container only - the reference does not travel to the GPU box).

This validates the stand-ins of oracle/refshim the only way available here: everything the reference's tests assert
about the behaviour of `cagpjax` on top of jax / CoLA / GPJax / lineax / numpyro / matfree must hold when those packages
are replaced by the stand-ins.  Seven of the eight test files run here (~15 s); `test_operators.py` (2 minutes, large
`LazyKernel` cases through the Python-loop `lax.map`) is run by tools/run_reference_tests_under_standins.sh and recorded
in profiles/r02_reference_suite_under_standins.md.  Tests that need automatic differentiation skip themselves (the
stand-in for jax has none).  Known failures are listed with their cause; anything else fails this test."""

import os
import re
import subprocess
import sys

import pytest

from oracle.run_reference import REFERENCE_SRC, SHIM_DIR, reference_available

REF_TESTS = os.path.join(os.path.dirname(REFERENCE_SRC), "tests")
FAST_FILES = ["test_computations.py", "test_interop.py", "test_solvers.py", "test_linalg.py", "test_policies.py",
              "test_distributions.py", "test_models.py"]
# The stand-in for jax.random draws from numpy's bit streams, not JAX's threefry, so three data-dependent cases see
# different inputs than upstream: a 2 x 2 `A A^T` that happens to be near-singular (the test compares a jittered
# Cholesky log-density with an unjittered one), and a Lanczos run whose start vector needs more than 8 iterations.
KNOWN_DATA_DEPENDENT = {
    "TestGaussianDistribution::test_log_prob_consistency_with_numpyro[float32-2-Cholesky]",
    "TestGaussianDistribution::test_log_prob_consistency_with_numpyro[float64-2-Cholesky]",
    "TestLanczosPolicy::test_eigenvectors_match_dense_computation[psd_linear_operator1-8]",
}


@pytest.mark.skipif(not (reference_available() and os.path.isdir(REF_TESTS)), reason="reference sources are not on this machine")
def test_reference_tests_pass_under_the_stand_ins(tmp_path):
    env = dict(os.environ, PYTEST_DISABLE_PLUGIN_AUTOLOAD="1", PYTHONPATH=os.pathsep.join([SHIM_DIR, REFERENCE_SRC]))
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-q", "-rf"]
    cmd += [os.path.join(REF_TESTS, f) for f in FAST_FILES]
    res = subprocess.run(cmd, env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    tail = res.stdout.strip().splitlines()[-1]
    failed = {line.split(" ", 1)[1].split(" - ")[0].split("::", 1)[1] for line in res.stdout.splitlines() if line.startswith("FAILED")}
    m = re.search(r"(\d+) passed", tail)
    assert m, res.stdout[-2000:] + res.stderr[-2000:]
    assert failed <= KNOWN_DATA_DEPENDENT, sorted(failed - KNOWN_DATA_DEPENDENT)
    assert "error" not in tail, tail
    assert int(m.group(1)) >= 2600, tail   # 2623 at the time of writing; 526 skip themselves (autodiff / float32 notes)
