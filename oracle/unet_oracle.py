"""CPU restatement of the reference's UNet forward — TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows models/network_unet.py:13-66 with the blocks of models/basicblock.py:61-63 (conv -> Conv2d/ConvTranspose2d, 'R' -> ReLU),
:413-419 (upsample_convtranspose: ConvTranspose2d kernel 2, stride 2, padding 0, + ReLU) and :439-445 (downsample_strideconv:
Conv2d kernel 2, stride 2, padding 0, + ReLU), written with torch.nn.functional on float64 tensors.

Pinned by tests/golden/unet.npz: outputs of the reference's OWN module graph.  The reference class cannot be constructed as
shipped (network_unet.py:17 calls load_state_dict(torch.load("")) before any layer exists), so tests/golden/make_golden.py
builds it with exactly that one statement neutralised (torch.load -> {} and load_state_dict -> no-op during __init__); every
layer, the forward and the parameter names are the reference's.
"""
from __future__ import annotations

import numpy as np


def unet_forward(state_dict, x0, nb=2):
    """state_dict: name -> ndarray (reference module names); x0: (C,H,W) or (B,C,H,W) -> float64 ndarray of the same shape."""
    import torch
    import torch.nn.functional as F
    P = lambda k: torch.from_numpy(np.asarray(state_dict[k], dtype=np.float64))
    x = torch.from_numpy(np.asarray(x0, dtype=np.float64))
    squeeze = x.dim() == 3
    if squeeze:
        x = x[None]

    def conv(t, key, relu=True):
        t = F.conv2d(t, P(key + ".weight"), P(key + ".bias"), stride=1, padding=1)
        return F.relu(t) if relu else t

    def down(t, key):
        return F.relu(F.conv2d(t, P(key + ".weight"), P(key + ".bias"), stride=2, padding=0))

    def up(t, key):
        return F.relu(F.conv_transpose2d(t, P(key + ".weight"), P(key + ".bias"), stride=2, padding=0))

    def level_down(t, name):                       # network_unet.py:31-33
        for k in range(nb):
            t = conv(t, f"{name}.{2 * k}")
        return down(t, f"{name}.{2 * nb}")

    def level_up(t, name):                         # network_unet.py:47-49
        t = up(t, f"{name}.0")
        for k in range(nb):
            t = conv(t, f"{name}.{2 * (k + 1)}")
        return t

    x1 = conv(x, "m_head.0")                       # network_unet.py:54
    x2 = level_down(x1, "m_down1")
    x3 = level_down(x2, "m_down2")
    x4 = level_down(x3, "m_down3")
    t = x4
    for k in range(nb + 1):                        # m_body, network_unet.py:35
        t = conv(t, f"m_body.{2 * k}")
    t = level_up(t + x4, "m_up3")                  # network_unet.py:59-61
    t = level_up(t + x3, "m_up2")
    t = level_up(t + x2, "m_up1")
    t = conv(t + x1, "m_tail", relu=False) + x     # network_unet.py:62
    out = t.numpy()
    return out[0] if squeeze else out
