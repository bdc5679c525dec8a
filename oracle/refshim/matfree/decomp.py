def tridiag_sym(num_matvecs, /, *, materialize=True, reortho="full", custom_vjp=True):
    raise NotImplementedError("matfree.decomp.tridiag_sym is not part of the stand-in (Lanczos is outside the pinned path)")
