"""The JAX-side drop-in (`jax_binding/cagpjax_b200_jax`) exercised END TO END on the CPU: the reference's own model runs
with `B200KernelComputation` as its compute engine under the dependency stand-ins of oracle/refshim, every XLA FFI target
is a numpy emulation of the C-ABI entry point's semantics, and the results must equal the reference's results with its
own `LazyKernelComputation` (the committed reference goldens).  See tests/jax_binding_wiring_driver.py for what this
does and does not check.  Needs the reference sources (this container; not the GPU box), gcc and the built library."""

import os
import re
import subprocess
import sys

import pytest

from oracle.run_reference import reference_available

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "cagpjax_b200", "_lib", "libcagp_b200.so")


@pytest.mark.skipif(not reference_available(), reason="reference sources are not on this machine")
@pytest.mark.skipif(not os.path.exists(LIB), reason="libcagp_b200.so is not built")
def test_reference_model_runs_on_the_jax_binding_with_emulated_targets(tmp_path):
    names = re.findall(r'"(Cagp[A-Za-z]+)":', open(os.path.join(ROOT, "jax_binding", "cagpjax_b200_jax", "ffi.py")).read())
    assert len(names) >= 15
    # a stand-in for libcagp_b200_xla.so (csrc/xla_ffi_shim.cc needs the XLA headers): the handler symbols only
    (tmp_path / "dummy.c").write_text("\n".join(f"void {n}(void) {{}}" for n in names))
    subprocess.run(["gcc", "-shared", "-fPIC", "-o", str(tmp_path / "libcagp_b200_xla.so"), str(tmp_path / "dummy.c")], check=True)
    os.symlink(LIB, tmp_path / "libcagp_b200.so")
    res = subprocess.run([sys.executable, os.path.join(HERE, "jax_binding_wiring_driver.py"), str(tmp_path),
                          os.path.join(HERE, "golden", "reference_golden.npz")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "WIRING OK" in res.stdout
