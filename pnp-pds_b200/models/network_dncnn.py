"""KAIR DnCNN with the reference's constructor (models/network_dncnn.py:42-77): x - model(x),
(conv+ReLU) x (nb-1) + conv, no clamps.  Used by the '*-unstable-*' methods (iteration.py:34-39)."""
from __future__ import annotations

import numpy as np

from .denoiser import Denoiser
from .weights import load_weights


class DnCNN:
    def __init__(self, in_nc=1, out_nc=1, nc=64, nb=17, act_mode="BR", model_path=""):
        assert "R" in act_mode or "L" in act_mode, "Examples of activation function: R, L, BR, BL, IR, IL"
        if "B" in act_mode:
            raise ValueError("batch-norm DnCNN variants are not supported (no such checkpoint ships with the reference)")
        w = load_weights(model_path)
        if (w.c_in, w.c_out, w.n_ch, w.depth) != (in_nc, out_nc, nc, nb):
            raise RuntimeError(f"checkpoint is {w.c_in}->{w.n_ch}x{w.depth}->{w.c_out}, requested {in_nc}->{nc}x{nb}->{out_nc}")
        self._den = Denoiser(w, in_nc)

    def __call__(self, x):
        """torch tensor (N,C,H,W) or (C,H,W) -> same shape (the reference passes an unbatched gray tensor, iteration.py:108)."""
        import torch
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))
        shape = t.shape
        t4 = t.reshape((-1,) + tuple(shape[-3:])) if t.dim() >= 3 else t.reshape(1, 1, *shape)
        out = self._den.denoise_batch(t4.float())
        return out.reshape(shape).to(t.device)

    forward = __call__

    def eval(self):
        return self
