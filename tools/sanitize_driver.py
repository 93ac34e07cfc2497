#!/usr/bin/env python
"""CI-size workload for compute-sanitizer (tools/gpu_sanitize.sh): every kernel family of libpnp_pds.so once or twice.

  body layers: row-streaming (conv_roll_d_kernel, conv_roll_kernel), CTA-pair tiles (conv_tc2_kernel), 1-CTA tiles
  (conv_tc_kernel), the chain kernel (conv_chain_kernel); first / last layer; blur stencils (apply, primal, dual, both tile
  shapes); pointwise primal / dual; l1 ball (register-cached and streaming); SSIM; TV.
Shapes are tiny on purpose: the sanitizer tools slow kernels down by 10-1000x.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    which = set(sys.argv[1:]) or {"dncnn", "loops", "l1", "tv"}
    import torch
    from pnp_pds_b200 import iteration, operators
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    h = np.load(os.path.join(GOLDEN, "assets.npz"))["blur_1"]
    w1 = load_weights(os.path.join(GOLDEN, "weights", "DnCNN_nobn_nch_1_nlev_0.01.pdsw"))
    w3 = load_weights(os.path.join(GOLDEN, "weights", "DnCNN_nobn_nch_3_nlev_0.01.pdsw"))
    rng = np.random.default_rng(0)
    if "dncnn" in which:
        for (B, C, H, W, variant, name) in [(2, 1, 20, 256, 64, "roll_d"), (1, 1, 20, 256, 64 | 256, "roll_hbm"), (2, 1, 40, 24, 128 | 512, "tc2"),
                                            (1, 3, 24, 24, 16 | 512, "tc1"), (2, 1, 40, 24, 128, "chain"), (1, 3, 33, 20, 128, "chain_c3")]:
            x = rng.random((B, C, H, W)).astype(np.float32)
            with Engine(B, C, H, W) as e:
                e.load_dncnn(w3 if C == 3 else w1)
                e.set_tc_variant(variant)
                for _ in range(2):
                    y = e.dncnn_forward(e.to_device(x))
                torch.cuda.synchronize()
                print(name, (B, C, H, W), float(y.mean()), flush=True)
    if "loops" in which:
        prmA = dict(gamma1=0.99, gamma2=0.99, alpha_s=0.95, alpha_n=0.95, myLambda=1.0, gaussian_nl=0.01, sp_nl=0.0, poisson_alpha=300, r=1.0)
        prmB = dict(gamma1=1.0, gamma2=0.49, alpha_s=0.9, alpha_n=0.9, myLambda=1.0, gaussian_nl=0.01, sp_nl=0.1, poisson_alpha=300, r=0.8)
        prmC = dict(gamma1=0.0006, gamma2=1 / 0.0006, alpha_s=0.95, alpha_n=0.9, myLambda=1.0, gaussian_nl=0.0, sp_nl=0.0, poisson_alpha=100, r=1.0)
        for (method, op, C, H, W, prm) in [("A-Proposed", "blur", 1, 40, 136, prmA), ("A-Proposed", "blur", 3, 32, 32, prmA),
                                           ("B-Proposed", "random_sampling", 1, 32, 40, prmB), ("B-Proposed", "blur", 1, 36, 36, prmB),
                                           ("C-Proposed", "blur", 1, 32, 32, prmC), ("A-Proposed", "Id", 1, 24, 24, prmA)]:
            phi, adj = operators.get_observation_operators(op, h, prm["r"])
            shape = (2, H, W) if C == 1 else (2, C, H, W)
            xt = (0.2 + 0.6 * rng.random(shape)).astype(np.float32)
            obs = np.stack([np.asarray(phi(z), dtype=np.float32) for z in xt])
            if method.startswith("C"):
                obs = np.round(obs * 100).astype(np.float32)
                x0 = obs / 100
            else:
                x0 = obs
            res = iteration.run_batch(x0, obs, xt, phi, adj, prm, w3 if C == 3 else w1, 3, method, C, ssim="all")
            print(method, op, shape, float(res["psnr"][-1, 0]), float(res["ssim"][-1, 0]), flush=True)
        iteration.clear_engine_cache()
    if "l1" in which:
        for n in (5000, 300000):                       # register-cached / streaming variant
            z = (rng.standard_normal(n) * 0.3).astype(np.float32)
            out = operators.proj_l1_ball(z, 0.9, 0.1, 0.8)
            print("l1ball", n, float(np.abs(out).sum()), flush=True)
        print("l2ball", float(np.linalg.norm(operators.proj_l2_ball(rng.random(4000), 0.9, 0.01, 0.0, rng.random(4000)))), flush=True)
        print("gkl", float(operators.prox_GKL(rng.random(4000) * 50, 1 / 1666.0, 100.0, rng.poisson(30, 4000)).sum()), flush=True)
    if "tv" in which:
        prm = dict(gamma1=0.1, gamma2=0.99, alpha_s=0.95, alpha_n=0.95, myLambda=1.0, gaussian_nl=0.01, sp_nl=0.0, poisson_alpha=300, r=1.0)
        phi, adj = operators.get_observation_operators("blur", h, 1.0)
        xt = (0.2 + 0.6 * rng.random((1, 3, 24, 28))).astype(np.float32)
        obs = np.stack([np.asarray(phi(z), dtype=np.float32) for z in xt])
        for m in ("A-PDS-TV", "A-FBS-TV", "comparisonB-3"):
            res = iteration.run_batch(obs, obs, xt, phi, adj, prm, None, 3, m, 3, ssim="none")
            print(m, float(res["psnr"][-1, 0]), flush=True)
        iteration.clear_engine_cache()


if __name__ == "__main__":
    main()
