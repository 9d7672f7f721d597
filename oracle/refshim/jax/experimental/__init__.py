"""``jax.experimental`` placeholder: the sparse BCSR/BCOO paths (spineax/cuDSS, dense-from-BCOO) are
not exercised by the golden generator, which uses the reference's dense and CSR-direct paths."""
import types

sparse = types.SimpleNamespace(BCSR=None, BCOO=None)
