"""KAIR UNet with the reference's constructor arguments (models/network_unet.py:13-66), forward on the B200.

The reference class cannot be constructed (it calls ``load_state_dict`` before any layer exists, network_unet.py:17) and
ships no weights, so this class takes the parameters as a ``state_dict`` (names as the reference's modules would give them:
``m_head.0.weight``, ``m_down1.{0,2,4}.*``, ``m_body.{0,2,4}.*``, ``m_up3.{0,2,4}.*``, ..., ``m_tail.weight``) or a file
``torch.save``d from one (tensors only, read with ``weights_only=True``).  Only the reference's default modes exist:
``act_mode='R'``, ``downsample_mode='strideconv'``, ``upsample_mode='convtranspose'``.  Compute: pds_unet_forward (csrc/unet.cu).
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from .. import _lib

_HDR = struct.Struct("<4s8i12x")     # magic, version, in_nc, out_nc, nc[4], nb  (48 bytes)


def layer_keys(nb: int = 2):
    """(state_dict prefix, kind) in module order: 'C' 3x3 conv, 'D' 2x2 stride-2 conv, 'T' 2x2 stride-2 transposed conv."""
    keys = [("m_head.0", "C")]
    for i in (1, 2, 3):
        keys += [(f"m_down{i}.{2 * k}", "C") for k in range(nb)] + [(f"m_down{i}.{2 * nb}", "D")]
    keys += [(f"m_body.{2 * k}", "C") for k in range(nb + 1)]
    for i in (3, 2, 1):
        keys += [(f"m_up{i}.0", "T")] + [(f"m_up{i}.{2 * (k + 1)}", "C") for k in range(nb)]
    keys.append(("m_tail", "C"))
    return keys


def layer_shapes(in_nc, out_nc, nc, nb):
    """Weight shape per layer in module order (torch layouts)."""
    shapes = [(nc[0], in_nc, 3, 3)]
    for i in range(3):
        shapes += [(nc[i], nc[i], 3, 3)] * nb + [(nc[i + 1], nc[i], 2, 2)]
    shapes += [(nc[3], nc[3], 3, 3)] * (nb + 1)
    for i in (3, 2, 1):
        shapes += [(nc[i], nc[i - 1], 2, 2)] + [(nc[i - 1], nc[i - 1], 3, 3)] * nb      # ConvTranspose2d: [cin][cout][k][k]
    shapes.append((out_nc, nc[0], 3, 3))
    return shapes


def random_state_dict(in_nc=1, out_nc=1, nc=(64, 128, 256, 512), nb=2, seed=0, scale=1.0):
    """He-style random parameters with the reference's names (there is no trained UNet to load)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for (key, kind), shp in zip(layer_keys(nb), layer_shapes(in_nc, out_nc, list(nc), nb)):
        fan_in = (shp[0] if kind == "T" else shp[1]) * shp[2] * shp[3]
        sd[key + ".weight"] = (rng.standard_normal(shp) * scale * np.sqrt(2.0 / fan_in)).astype(np.float32)
        sd[key + ".bias"] = (rng.standard_normal(shp[1] if kind == "T" else shp[0]) * 0.05).astype(np.float32)
    return sd


def to_blob(state_dict, in_nc, out_nc, nc, nb) -> bytes:
    out = [_HDR.pack(b"PDSU", 1, in_nc, out_nc, *[int(c) for c in nc], nb)]
    for (key, kind), shp in zip(layer_keys(nb), layer_shapes(in_nc, out_nc, list(nc), nb)):
        w = np.asarray(_np(state_dict[key + ".weight"]), dtype="<f4")
        b = np.asarray(_np(state_dict[key + ".bias"]), dtype="<f4")
        if tuple(w.shape) != tuple(shp) or b.shape != ((shp[1] if kind == "T" else shp[0]),):
            raise ValueError(f"{key}: expected weight {shp}, got {tuple(w.shape)}")
        out += [np.ascontiguousarray(w).tobytes(), np.ascontiguousarray(b).tobytes()]
    return b"".join(out)


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


class UNet:
    def __init__(self, file_name="", in_nc=1, out_nc=1, nc=(64, 128, 256, 512), nb=2, act_mode="R", downsample_mode="strideconv",
                 upsample_mode="convtranspose", state_dict=None):
        if act_mode != "R" or downsample_mode != "strideconv" or upsample_mode != "convtranspose":
            raise NotImplementedError("only the reference's defaults: act_mode='R', strideconv, convtranspose (network_unet.py:14)")
        if in_nc != out_nc:
            raise ValueError("the input residual (network_unet.py:62) needs out_nc == in_nc")
        if state_dict is None:
            if not file_name:
                raise ValueError("UNet needs a state_dict or a file saved from one (the reference ships no UNet weights)")
            import torch
            state_dict = torch.load(file_name, map_location="cpu", weights_only=True)
        self.in_nc, self.out_nc, self.nc, self.nb = int(in_nc), int(out_nc), [int(c) for c in nc], int(nb)
        self._blob = to_blob(state_dict, self.in_nc, self.out_nc, self.nc, self.nb)
        self._handles = {}                     # (B, H, W) -> pds_unet_t

    def _handle(self, B, H, W):
        import torch
        key = (B, H, W)
        h = self._handles.get(key)
        if h is None:
            if not torch.cuda.is_available():
                raise _lib.PdsError("no CUDA device visible: pnp_pds_b200 has no CPU fallback")
            lib = _lib.load()
            cfg = _lib.PdsUnetConfig(B, self.in_nc, self.out_nc, (C.c_int32 * 4)(*self.nc), self.nb, H, W, torch.cuda.current_device())
            h = C.c_void_p()
            _lib.check(lib.pds_unet_create(C.byref(cfg), C.byref(h)))
            if lib.pds_unet_blob_bytes(C.byref(cfg)) != len(self._blob):
                lib.pds_unet_destroy(h)
                raise RuntimeError("UNet blob size does not match the configuration")
            try:
                _lib.check(lib.pds_unet_load(h, self._blob, len(self._blob)))
            except Exception:
                lib.pds_unet_destroy(h)
                raise
            while len(self._handles) >= 4:
                lib.pds_unet_destroy(self._handles.pop(next(iter(self._handles))))
            self._handles[key] = h
        return h

    def forward(self, x0):
        """numpy / torch (C,H,W) or (B,C,H,W) -> same kind and shape, float32 (network_unet.py:52-64)."""
        import torch
        is_t = isinstance(x0, torch.Tensor)
        t = x0 if is_t else torch.from_numpy(np.ascontiguousarray(np.asarray(x0), dtype=np.float32))
        shape = t.shape
        t4 = t.reshape((1,) + tuple(shape)) if t.dim() == 3 else t
        if t4.dim() != 4 or t4.shape[1] != self.in_nc:
            raise ValueError(f"expected (B,{self.in_nc},H,W) or ({self.in_nc},H,W)")
        B, _, H, W = t4.shape
        h = self._handle(B, H, W)
        dev = torch.device("cuda", torch.cuda.current_device())
        xd = t4.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty_like(xd)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.load().pds_unet_forward(h, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), st))
        out = out.reshape(shape)
        return out.to(t.device) if is_t else out.cpu().numpy()

    __call__ = forward

    def eval(self):
        return self

    def close(self):
        lib = _lib.load()
        while self._handles:
            lib.pds_unet_destroy(self._handles.popitem()[1])

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
