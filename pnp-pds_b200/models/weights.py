"""DnCNN checkpoint loading and the flat weight blob handed to the C-ABI.

Reference: models/denoiser.py:18-32 (pickled ``DataParallel(simple_CNN)`` checkpoints,
``checkpoint.module.state_dict()``) and models/network_dncnn.py:71 (plain KAIR
``state_dict`` with keys ``model.{0,2,..}.{weight,bias}``).

The reference's ``.pth`` files are legacy torch pickles of whole modules.  They are
untrusted input, so they are read with an allow-listed unpickler (only the ten
globals the six shipped checkpoints use) and immediately flattened to a ``.pdsw``
blob: a 48-byte header followed by fp32 ``weight[Cout][Cin][3][3]`` and
``bias[Cout]`` per layer.  ``pds_load_dncnn`` consumes that blob.
"""
from __future__ import annotations

import collections
import io
import os
import pickle
import struct
import types
from dataclasses import dataclass

import numpy as np

MAGIC = b"PDSW"
_HDR = struct.Struct("<4siiiiiffii8x")  # magic, version, depth, c_in, n_ch, c_out, slope, res_sign, clamp, reserved


@dataclass
class DnCNNWeights:
    """Flat description of a 3x3-conv chain: conv+act x (depth-1), conv, residual."""
    layers: list            # [(w (Cout,Cin,3,3) f32, b (Cout,) f32)]
    slope: float            # LeakyReLU negative slope (0.01 simple_CNN, 0.0 KAIR ReLU)
    residual_sign: float    # +1: out = net(x) + x   (basic_models.py:36);  -1: out = x - net(x) (network_dncnn.py:77)
    clamp: bool             # clamp input and output to [0,1] (denoiser.py:40-42)

    @property
    def depth(self):
        return len(self.layers)

    @property
    def c_in(self):
        return int(self.layers[0][0].shape[1])

    @property
    def n_ch(self):
        return int(self.layers[0][0].shape[0])

    @property
    def c_out(self):
        return int(self.layers[-1][0].shape[0])

    def to_blob(self) -> bytes:
        out = [_HDR.pack(MAGIC, 1, self.depth, self.c_in, self.n_ch, self.c_out, float(self.slope),
                         float(self.residual_sign), int(self.clamp), 0)]
        for w, b in self.layers:
            out.append(np.ascontiguousarray(w, dtype="<f4").tobytes())
            out.append(np.ascontiguousarray(b, dtype="<f4").tobytes())
        return b"".join(out)

    @staticmethod
    def from_blob(buf: bytes) -> "DnCNNWeights":
        magic, ver, depth, c_in, n_ch, c_out, slope, rs, clamp, _ = _HDR.unpack_from(buf, 0)
        if magic != MAGIC or ver != 1:
            raise ValueError("not a PDSW v1 weight blob")
        off = _HDR.size
        layers = []
        for i in range(depth):
            ci = c_in if i == 0 else n_ch
            co = c_out if i == depth - 1 else n_ch
            nw = co * ci * 9
            w = np.frombuffer(buf, dtype="<f4", count=nw, offset=off).reshape(co, ci, 3, 3).copy()
            off += nw * 4
            b = np.frombuffer(buf, dtype="<f4", count=co, offset=off).copy()
            off += co * 4
            layers.append((w, b))
        if off != len(buf):
            raise ValueError("PDSW blob has trailing or missing bytes")
        return DnCNNWeights(layers, slope, rs, bool(clamp))


def _restricted_pickle_module():
    import torch
    import torch.nn as nn

    class simple_CNN(nn.Module):  # stand-in for models.basic_models.simple_CNN; only holds state
        pass

    allowed = {
        ("models.basic_models", "simple_CNN"): simple_CNN,
        ("collections", "OrderedDict"): collections.OrderedDict,
        ("torch.nn.parallel.data_parallel", "DataParallel"): nn.DataParallel,
        ("torch.nn.modules.conv", "Conv2d"): nn.Conv2d,
        ("torch.nn.modules.container", "ModuleList"): nn.ModuleList,
        ("torch.nn.modules.activation", "LeakyReLU"): nn.LeakyReLU,
        ("torch._utils", "_rebuild_tensor_v2"): torch._utils._rebuild_tensor_v2,
        ("torch._utils", "_rebuild_parameter"): torch._utils._rebuild_parameter,
        ("torch", "FloatStorage"): torch.FloatStorage,
        ("torch", "device"): torch.device,
    }

    class RestrictedUnpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return allowed[(module, name)]
            except KeyError:
                raise pickle.UnpicklingError(f"checkpoint references a global outside the allow-list: {module}.{name}")

    # torch's legacy (non-zip) loader also calls pickle_module.load(f) for the magic number, protocol version, sys-info and the
    # storage keys, and the tar branch calls it too: every entry point goes through the allow-listed unpickler.
    def _load(f, **kw):
        return RestrictedUnpickler(f, **kw).load()

    def _loads(b, **kw):
        return RestrictedUnpickler(io.BytesIO(b), **kw).load()

    return types.SimpleNamespace(Unpickler=RestrictedUnpickler, load=_load, loads=_loads,
                                 __name__="pnp_pds_b200.restricted_pickle")


def load_pth(path: str) -> DnCNNWeights:
    """Read one of the reference's ``nn/*.pth`` files (either flavour)."""
    import torch

    obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_restricted_pickle_module())
    if isinstance(obj, dict):  # KAIR DnCNN state_dict: model.{0,2,...}
        idx = sorted({int(k.split(".")[1]) for k in obj})
        layers = [(obj[f"model.{i}.weight"].float().numpy().copy(), obj[f"model.{i}.bias"].float().numpy().copy())
                  for i in idx]
        return DnCNNWeights(layers, slope=0.0, residual_sign=-1.0, clamp=False)
    mod = obj.module if hasattr(obj, "module") else obj
    sd = mod.state_dict()
    if getattr(mod, "bn", False):
        raise ValueError("batch-norm simple_CNN checkpoints are not supported (none ship with the reference)")
    n_mid = len({k.split(".")[1] for k in sd if k.startswith("conv_list.")})
    names = ["in_conv"] + [f"conv_list.{i}" for i in range(n_mid)] + ["out_conv"]
    layers = [(sd[f"{n}.weight"].float().numpy().copy(), sd[f"{n}.bias"].float().numpy().copy()) for n in names]
    slopes = {float(m.negative_slope) for m in mod.nl_list}
    if len(slopes) != 1:
        raise ValueError("mixed activation slopes")
    return DnCNNWeights(layers, slope=slopes.pop(), residual_sign=1.0, clamp=True)


_LOADED: dict = {}      # (realpath, mtime_ns, size) -> DnCNNWeights: a checkpoint is parsed once per process, not once per call


def load_weights(path: str) -> DnCNNWeights:
    """Load ``path`` (.pth or .pdsw).  If ``path`` does not exist but a sibling
    ``<stem>.pdsw`` does, that is used (the GPU box only carries converted blobs)."""
    if os.path.exists(path):
        st = os.stat(path)
        key = (os.path.realpath(path), st.st_mtime_ns, st.st_size)
        w = _LOADED.get(key)
        if w is None:
            if path.endswith(".pdsw"):
                with open(path, "rb") as f:
                    w = DnCNNWeights.from_blob(f.read())
            else:
                w = load_pth(path)
            if len(_LOADED) >= 8:
                _LOADED.pop(next(iter(_LOADED)))
            _LOADED[key] = w
        return w
    alt = os.path.splitext(path)[0] + ".pdsw"
    if os.path.exists(alt):
        return load_weights(alt)
    raise FileNotFoundError(path)


def convert(path_pth: str, path_out: str) -> None:
    w = load_pth(path_pth)
    with open(path_out, "wb") as f:
        f.write(w.to_blob())
