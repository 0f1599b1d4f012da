"""fairguide: B200-native (sm_100a) kernels for the fairness-guidance path of
sail-sg/finetune-fair-diffusion, behind the Python call surface of 1-main-debias.py.

The directory name carries the upstream repo name (with a hyphen, so it is not importable by
name); import it as ``fairguide`` (see ../fairguide/__init__.py).
"""
from . import _lib, api, autograd, dist, ops, sync  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import bind  # noqa: F401
from .sync import GradBucket, adjusted_dft_grad_coefs, allreduce_average_gradients, make_grad_hook  # noqa: F401

__all__ = [n for n in dir(api) if not n.startswith("_")] + ["GradBucket", "adjusted_dft_grad_coefs", "allreduce_average_gradients", "make_grad_hook"]
