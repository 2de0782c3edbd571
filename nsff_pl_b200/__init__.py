"""nsff_pl_b200 - Blackwell-native (sm_100a) implementation of the NSFF render_rays hot path.

Drop-in surface (same names and signatures as kwea123/nsff_pl):
    from nsff_pl_b200.models.nerf import PosEmbedding, NeRF
    from nsff_pl_b200.models.rendering import render_rays, sample_pdf
or `nsff_pl_b200.install_as_reference_modules()` to serve them as `models.nerf` / `models.rendering`
so the reference's train.py / eval.py import them unchanged.

The compute lives in libnsff_b200.so (hand-written CUDA behind the C ABI in include/nsff_b200.h).
There is no CPU path: calling any op without the built library or with CPU tensors raises.
"""
import sys
import types

__version__ = "0.1.0"


def install_as_reference_modules():
    """Register this package's modules under the reference's import paths (`models.nerf`,
    `models.rendering`) so `from models.rendering import render_rays` in train.py/eval.py resolves here."""
    from .models import nerf, rendering
    pkg = sys.modules.get("models")
    if pkg is None:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
    pkg.nerf, pkg.rendering = nerf, rendering
    sys.modules["models.nerf"] = nerf
    sys.modules["models.rendering"] = rendering
    return pkg
