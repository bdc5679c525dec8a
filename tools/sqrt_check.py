"""Accuracy of the Goldschmidt square root of csrc/kernel_math.cuh (fast_sqrt), emulated with exact FMAs: seeds off by
up to 2^-20 relative (MUFU.RSQ64H is ~2^-22) still give <= 2.1 ulp.  CPU only."""
import numpy as np

rng = np.random.default_rng(0)
x = 10 ** rng.uniform(-12, 3, 200000)
ld = np.longdouble
fma = lambda a, b, c: np.float64(ld(a) * ld(b) + ld(c))
for delta in (2**-22, -2**-22, 2**-20, -2**-20):
    y = np.float64((1 / np.sqrt(x)) * (1 + delta))
    g, h = x * y, 0.5 * y
    e = fma(-g, h, 0.5); g = fma(g, e, g); h = fma(h, e, h)
    e = fma(-g, h, 0.5); g = fma(g, e, g)
    ref = np.sqrt(ld(x))
    err = float(np.max(np.abs((ld(g) - ref) / ref)))
    print(f"seed error {delta:+.1e}: max relative error {err:.3e} = {err / 2**-53:.2f} ulp")
