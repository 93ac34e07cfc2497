"""Denoiser wrapper with the reference's interface (models/denoiser.py:9-46).

Denoiser(file_name, ch).denoise(x): numpy (H,W) | (C,H,W), any float  ->  numpy float32 same shape,
= clamp(net(clamp(x,0,1)),0,1) with net = simple_CNN (basic_models.py:25-38), computed by the
conv_first / tcgen05 mid / conv_last kernels.
"""
from __future__ import annotations

import numpy as np

from ..engine import Engine
from .weights import DnCNNWeights, load_weights

_cache: dict = {}


class Denoiser:
    def __init__(self, file_name, ch, conv_engine: str = "tcgen05"):
        self.cost = 0
        self.ch = int(ch)
        self.file_name = file_name
        self.conv_engine = conv_engine
        self.weights: DnCNNWeights = file_name if isinstance(file_name, DnCNNWeights) else load_weights(str(file_name))
        if self.weights.c_in != self.ch:
            raise ValueError(f"checkpoint has {self.weights.c_in} channels, ch={ch}")
        self._engines: dict = {}

    @classmethod
    def cached(cls, file_name, ch):
        key = (str(file_name), int(ch))
        if key not in _cache:
            _cache[key] = cls(file_name, ch)
        return _cache[key]

    def engine(self, B: int, H: int, W: int) -> Engine:
        key = (B, H, W)
        e = self._engines.get(key)
        if e is None:
            e = Engine(B, self.ch, H, W, method="A", deg_op="Id", max_iter=1, conv_engine=self.conv_engine)
            e.load_dncnn(self.weights)
            if len(self._engines) > 8:
                self._engines.pop(next(iter(self._engines))).close()
            self._engines[key] = e
        return e

    def denoise(self, x):
        x = np.asarray(x)
        if x.ndim == 2:
            H, W = x.shape
        elif x.ndim == 3 and x.shape[0] == self.ch:
            _, H, W = x.shape
        else:
            raise ValueError(f"expected (H,W) or ({self.ch},H,W), got {x.shape}")
        e = self.engine(1, H, W)
        out = e.dncnn_forward(e.to_device(x))
        return out.cpu().numpy().reshape(x.shape)

    def denoise_batch(self, x):
        """(B,C,H,W) numpy or device tensor -> same kind."""
        import torch
        B, C, H, W = x.shape
        e = self.engine(B, H, W)
        out = e.dncnn_forward(e.to_device(x))
        return out if isinstance(x, torch.Tensor) else out.cpu().numpy()
