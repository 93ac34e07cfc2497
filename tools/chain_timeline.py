#!/usr/bin/env python
"""Timeline of the chain kernel's pipeline (pds_debug_chain_trace): where a unit's time goes, per role.
usage: python tools/chain_timeline.py [H W [tc_variant]]   (one gray image, default 256 256, variant 128)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EV = ["polled", "tma", "tempty", "full0", "issued", "tfull", "stored", "published"]


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 256)
    variant = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    import torch
    from pnp_pds_b200 import _lib
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    w = load_weights(os.path.join(ROOT, "tests", "golden", "weights", "DnCNN_nobn_nch_1_nlev_0.01.pdsw"))
    x = np.random.default_rng(0).random((1, 1, H, W)).astype(np.float32)
    with Engine(1, 1, H, W) as e:
        e.load_dncnn(w)
        e.set_tc_variant(variant)
        xd = e.to_device(x)
        for _ in range(3):
            e.dncnn_forward(xd)
        torch.cuda.synchronize()
        _lib.check(e.lib.pds_debug_chain_trace(e._h, None))
        e.dncnn_forward(xd)
        torch.cuda.synchronize()
        out = np.zeros((4, 64, 8), dtype=np.uint64)
        _lib.check(e.lib.pds_debug_chain_trace(e._h, out.ctypes.data_as(C.c_void_p)))
    t = out.astype(np.float64)
    t0 = t[t > 0].min()
    for c in range(2):
        t0 = t[c][t[c] > 0].min()
        print(f"cluster {c}: SM cycles / 1000 since the cluster's first stamp; per unit k")
        print("  k " + " ".join(f"{n:>9s}" for n in EV) + "   | issue-issue  tfull->stored  stored->pub")
        prev_issue = None
        for k in range(40):
            row = t[c, k]
            if row[4] == 0:
                break
            r = [(v - t0) / 1e3 if v > 0 else float("nan") for v in row]
            d = (r[4] - prev_issue) if prev_issue is not None else float("nan")
            prev_issue = r[4]
            print(f"{k:3d} " + " ".join(f"{v:9.2f}" for v in r) + f"   | {d:9.2f} {r[6] - r[5]:9.2f} {r[7] - r[6]:9.2f}")
    iss = t[:, :, 4]
    for c in range(4):
        v = iss[c][iss[c] > 0]
        if len(v) > 8:
            print(f"cluster {c}: mean unit period {np.mean(np.diff(v[4:])):.0f} cycles over {len(v) - 5} units")


if __name__ == "__main__":
    main()
