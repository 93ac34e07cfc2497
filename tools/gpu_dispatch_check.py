#!/usr/bin/env python
"""Times one denoiser forward with the body layers forced onto the row-streaming kernels / the tile kernels / the
library's own dispatch, for a few launch shapes: the calibration check of the cost model in dncnn_roll.cu (roll_band_rows)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pnp_pds_b200 import _lib  # noqa: E402
from pnp_pds_b200.engine import Engine  # noqa: E402
from pnp_pds_b200.models.weights import load_weights  # noqa: E402

w = load_weights(os.path.join(ROOT, "tests", "golden", "weights", "DnCNN_nobn_nch_1_nlev_0.01.pdsw"))
lib = _lib.load()
shapes = [(16, 321, 481), (16, 481, 321), (68, 321, 481), (16, 200, 130), (4, 180, 180), (12, 256, 256), (1, 512, 512), (1, 512, 384),
          (1, 768, 512), (2, 256, 256), (4, 256, 256)]
for B, H, W in shapes:
    x = np.random.default_rng(0).random((B, 1, H, W)).astype(np.float32)
    res = {}
    for name, variant in (("roll", 64), ("tile", 128), ("auto", 0)):
        if name == "roll" and lib.pds_debug_roll_band_rows(B, H, W, 1) == 0:
            res[name] = float("nan")
            continue
        with Engine(B, 1, H, W, conv_engine="tcgen05") as e:
            e.load_dncnn(w)
            e.set_tc_variant(variant)
            xd = e.to_device(x)
            for _ in range(3):
                e.dncnn_forward(xd)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                e.dncnn_forward(xd)
            t1.record()
            torch.cuda.synchronize()
            res[name] = t0.elapsed_time(t1) / 10
    pick = lib.pds_debug_roll_band_rows(B, H, W, 0)
    best = "roll" if res["roll"] < res["tile"] else "tile"
    print(f"{B:3d} x {H}x{W}: roll {res['roll']:.3f} ms  tile {res['tile']:.3f} ms  auto {res['auto']:.3f} ms  model picks {'roll' if pick else 'tile'}"
          f"  {'OK' if (pick > 0) == (best == 'roll') else 'MISPICK %.0f%%' % (100 * abs(res['roll'] - res['tile']) / min(res['roll'], res['tile']))}")
