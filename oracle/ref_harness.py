"""Import the UNMODIFIED reference (``/root/reference``) for golden-vector generation.

TEST INFRASTRUCTURE ONLY; works only where the reference tree is mounted (the build
container — never the GPU box).  Four shims, as listed in SURVEY.md §8c:

1. ``bm3d`` (imported at the top of iteration.py:3, wheel absent) -> empty stub module.
2. ``skimage.metrics.structural_similarity`` (utils_eval.py:2, package absent) -> stub
   returning 0.0  ==> SSIM parity is UNPINNED.
3. ``torch.load`` defaults to ``weights_only=True`` since torch 2.6 and rejects the
   pickled ``DataParallel`` checkpoints -> default flipped to False for the reference.
4. ``torch.cuda.synchronize()`` (iteration.py:192) raises without a GPU -> no-op.

Nothing here copies reference source; it only arranges ``sys.modules`` so that the
reference's own files import and run.
"""
from __future__ import annotations

import functools
import os
import sys
import types

REF_ROOT = os.environ.get("PDS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "iteration.py"))


@functools.lru_cache(maxsize=1)
def load():
    """Returns a namespace with the reference modules: operators, iteration, utils_noise, utils_eval, admm."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    if "bm3d" not in sys.modules:
        sys.modules["bm3d"] = types.ModuleType("bm3d")
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        skm = types.ModuleType("skimage.metrics")
        skm.structural_similarity = lambda **kw: 0.0
        sk.metrics = skm
        sys.modules["skimage"] = sk
        sys.modules["skimage.metrics"] = skm
    if not getattr(torch.load, "_pds_shim", False):
        _orig = torch.load

        def _load(*a, **kw):
            kw.setdefault("weights_only", False)
            return _orig(*a, **kw)

        _load._pds_shim = True
        torch.load = _load
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **kw: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import operators as ref_operators          # noqa: E402
    import iteration as ref_iteration          # noqa: E402
    from utils import utils_noise, utils_eval  # noqa: E402
    from algorithm import admm                 # noqa: E402
    from models import denoiser as ref_denoiser

    return types.SimpleNamespace(operators=ref_operators, iteration=ref_iteration, utils_noise=utils_noise,
                                 utils_eval=utils_eval, admm=admm, denoiser=ref_denoiser, root=REF_ROOT)
