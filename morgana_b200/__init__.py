"""morgana_b200 -- B200-native (sm_100a) kernels behind morgana's per-batch frame-rate feature API.

Importing this package loads ``morgana_b200/lib/libmorgana_b200.so`` (build it with ``python -m morgana_b200.build``);
there is no CPU fallback, so a missing library is an ImportError and non-CUDA tensors are a RuntimeError.
"""
import sys as _sys

# `python -m morgana_b200.build` imports this package first; the library it is about to (re)build must not be required.
_building = 'morgana_b200.build' in getattr(_sys, 'orig_argv', [])

if not _building:
    from morgana_b200 import _lib          # noqa: F401  (loads the shared library or raises)
    from morgana_b200 import ops           # noqa: F401
    from morgana_b200 import utils         # noqa: F401
    from morgana_b200 import losses        # noqa: F401
    from morgana_b200 import metrics       # noqa: F401
    from morgana_b200 import data          # noqa: F401
    from morgana_b200 import nn            # noqa: F401
    from morgana_b200 import torch_ops     # noqa: F401  (registers torch.ops.morgana_b200.*)
    from morgana_b200.patch import patch, unpatch   # noqa: F401

__version__ = '0.1.0'
