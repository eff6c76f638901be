"""femb200 -- B200-native hot path of the mechanic2d elasticity examples of
SalzmanA/fem-libraries: batched element tangents, CSR pattern + write-once
assembly, assembled / matrix-free operator apply, (Jacobi-)PCG, 1-8 GPUs.

Python host over the C ABI of include/femb200.h (libfemb200.so, hand-written
sm_100a CUDA).  No CPU fallback."""
from . import _capi as capi  # noqa: F401
from . import mesh  # noqa: F401
from .mesh import P1, P2, Q2, Mesh  # noqa: F401


def __getattr__(name):
    # fem / dist import torch lazily so that `import femb200` stays cheap
    if name in ("fem", "dist"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
