#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run once in the build container:  python tests/golden/make_golden.py [--long]
The reference cannot travel to the GPU box, so its inputs/outputs are frozen here.
numpy / torch versions are stored in every file (numpy-2 FFT of float32 is complex64,
SURVEY.md §7).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from oracle.pds_oracle import synthetic_image  # noqa: E402

KERNEL = os.path.join(ref_harness.REF_ROOT, "blur_models", "blur_1.mat")
NN = os.path.join(ref_harness.REF_ROOT, "nn")


def versions():
    import torch
    return dict(numpy_version=np.__version__, torch_version=torch.__version__)


def save(name, **arrs):
    arrs.update(versions())
    np.savez_compressed(os.path.join(HERE, name), **arrs)
    sz = os.path.getsize(os.path.join(HERE, name))
    print(f"wrote {name}: {len(arrs)} arrays, {sz/1024:.1f} KiB")


def gen_assets(ref):
    import scipy.io
    b1 = scipy.io.loadmat(KERNEL)["blur"]
    b2 = scipy.io.loadmat(os.path.join(ref.root, "blur_models", "square_mini.mat"))["blur"]
    save("assets.npz", blur_1=b1, square_mini=b2)


def gen_ops(ref):
    op = ref.operators
    out = {}
    rng = np.random.default_rng(7)
    phi, adj = op.get_observation_operators("blur", KERNEL, 0.8)
    for tag, shape in (("g32", (32, 32)), ("c24", (3, 24, 40)), ("g48x20", (48, 20))):
        x64 = rng.random(shape)
        out[f"blur_{tag}_x"] = x64
        out[f"blur_{tag}_phi"] = phi(x64)
        out[f"blur_{tag}_adj"] = adj(x64)
        x32 = x64.astype(np.float32)
        out[f"blur_{tag}_phi_f32in"] = phi(x32)
        out[f"blur_{tag}_adj_f32in"] = adj(x32)
    # random-sampling masks (integer work: bit-exact)
    for (H, W, r) in ((32, 32, 0.8), (64, 64, 0.5), (48, 20, 0.7), (256, 256, 0.8), (512, 512, 0.8), (1024, 1024, 0.8)):
        phi_r, _ = op.get_observation_operators("random_sampling", KERNEL, r)
        m = phi_r(np.ones((H, W)))
        assert set(np.unique(m)) <= {0.0, 1.0}
        out[f"mask_{H}_{W}_{r}"] = np.packbits(m.astype(np.uint8).reshape(-1))
    phi_r, _ = op.get_observation_operators("random_sampling", KERNEL, 0.8)
    xc = rng.random((3, 16, 16))
    out["rs_c16_x"] = xc
    out["rs_c16_out"] = phi_r(xc)
    xg = rng.random((16, 16)).astype(np.float32)
    out["rs_g16_x"] = xg
    out["rs_g16_out"] = phi_r(xg)
    # prox / projections
    x = rng.standard_normal((3, 20, 20)) * 0.3
    b = rng.random((3, 20, 20))
    out["l2_x"], out["l2_b"] = x, b
    out["l2_params"] = np.array([0.9, 0.05, 0.1, 0.8])      # alpha_n, gaussian_nl, sp_nl, r
    out["l2_out"] = op.proj_l2_ball(x, 0.9, 0.05, 0.1, b, 0.8)
    out["l2_out_inside"] = op.proj_l2_ball(b + 1e-4 * x, 0.9, 0.05, 0.1, b, 0.8)
    z = rng.standard_normal((3, 20, 20)) * 0.4
    out["l1_x"] = z
    out["l1_params"] = np.array([0.9, 0.1, 0.8])            # alpha_s, sp_nl, r
    out["l1_out"] = op.proj_l1_ball(z, 0.9, 0.1, 0.8)
    out["l1_out_inside"] = op.proj_l1_ball(z * 1e-3, 0.9, 0.1, 0.8)
    zz = rng.standard_normal((1, 64, 64)) * 0.2
    zz[0, ::7, ::5] += 1.0
    out["l1b_x"] = zz
    out["l1b_out"] = op.proj_l1_ball(zz, 0.9, 0.1, 0.8)
    xg = rng.standard_normal((24, 24)) * 50
    x0 = rng.poisson(30, size=(24, 24)).astype(np.int64)
    out["gkl_x"], out["gkl_x0"] = xg, x0
    out["gkl_params"] = np.array([1.0 / 1666.0, 100.0])     # gamma, alpha
    out["gkl_out"] = op.prox_GKL(xg, 1.0 / 1666.0, 100.0, x0)
    save("ops.npz", **out)


def gen_noise(ref):
    un, op = ref.utils_noise, ref.operators
    out = {}
    ident, _ = op.get_observation_operators("Id", KERNEL, 0.8)
    rs, _ = op.get_observation_operators("random_sampling", KERNEL, 0.8)
    img = synthetic_image(0, 1, 64, 64).astype(np.float64)
    imgc = synthetic_image(1, 3, 32, 32).astype(np.float64)
    out["img_g"], out["img_c"] = img, imgc
    out["gauss_g_id"] = un.add_gaussian_noise(img, 0.01, ident)
    out["gauss_g_rs"] = un.add_gaussian_noise(rs(img), 0.01, rs)
    out["gauss_c_id"] = un.add_gaussian_noise(imgc, 0.02, ident)
    out["poisson_g"] = un.apply_poisson_noise(img, 100)
    out["sp_g_id"] = un.add_salt_and_pepper_noise(img, 0.1, ident)
    out["sp_g_rs"] = un.add_salt_and_pepper_noise(rs(img), 0.1, rs)
    out["sp_c_rs"] = un.add_salt_and_pepper_noise(rs(imgc), 0.1, rs)
    out["sp_g_zero"] = un.add_salt_and_pepper_noise(img, 0.0, ident)
    save("noise.npz", **out)


def gen_denoiser(ref):
    import torch
    out = {}
    rng = np.random.default_rng(11)
    for arch, ch in (("DnCNN_nobn_nch_1_nlev_0.01", 1), ("DnCNN_nobn_nch_3_nlev_0.01", 3), ("DnCNN_nobn_nch_1_nlev_0.009", 1)):
        den = ref.denoiser.Denoiser(file_name=os.path.join(NN, arch + ".pth"), ch=ch)
        shape = (40, 24) if ch == 1 else (3, 24, 40)
        x = synthetic_image(3, ch, shape[-2], shape[-1]).astype(np.float64) + 0.05 * rng.standard_normal(shape)
        x[..., 0, 0] = -0.3
        x[..., 1, 1] = 1.4          # exercise the input clamp
        out[f"{arch}_x"] = x
        out[f"{arch}_y"] = den.denoise(np.copy(x))
    from models.network_dncnn import DnCNN as KAIR
    for arch, ch, nb in (("dncnn_15", 1, 17), ("dncnn_color_blind", 3, 20), ("dncnn3", 1, 20)):
        net = KAIR(in_nc=ch, out_nc=ch, nc=64, nb=nb, act_mode="R", model_path=os.path.join(NN, arch + ".pth"))
        shape = (1, 1, 24, 40) if ch == 1 else (1, 3, 24, 40)
        x = (synthetic_image(4, ch, 24, 40).reshape(shape[1:]) + 0.05 * rng.standard_normal(shape[1:])).astype(np.float32)
        with torch.no_grad():
            y = net(torch.from_numpy(x)[None]).numpy()[0]
        out[f"{arch}_x"] = x
        out[f"{arch}_y"] = y
    save("denoiser.npz", **out)


LOOP_CASES = [
    # tag, method, deg_op, ch, (H,W), settings
    dict(tag="A_blur_g", method="A-Proposed", deg_op="blur", ch=1, hw=(64, 64), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="A_rs_g", method="A-Proposed", deg_op="random_sampling", ch=1, hw=(64, 64), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=0.8, iters=30),
    dict(tag="A_id_g", method="A-Proposed", deg_op="Id", ch=1, hw=(48, 40), gaussian_nl=0.02, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="A_blur_c", method="A-Proposed", deg_op="blur", ch=3, hw=(32, 48), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="B_rs_g", method="B-Proposed", deg_op="random_sampling", ch=1, hw=(64, 64), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=1.0, gamma2=0.49, alpha_n=0.9, alpha_s=0.9, r=0.8, iters=30),
    dict(tag="B_blur_g", method="B-Proposed", deg_op="blur", ch=1, hw=(64, 64), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=1.0, gamma2=0.49, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="B_rs_c", method="B-Proposed", deg_op="random_sampling", ch=3, hw=(32, 32), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=1.0, gamma2=0.49, alpha_n=0.9, alpha_s=0.9, r=0.8, iters=30),
    dict(tag="C_blur_g", method="C-Proposed", deg_op="blur", ch=1, hw=(64, 64), gaussian_nl=0.0, sp_nl=0.0,
         poisson_noise=True, poisson_alpha=100, gamma1=0.0006, gamma2=1 / 0.0006, myLambda=1.0,
         alpha_n=0.9, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="C_id_g", method="C-Proposed", deg_op="Id", ch=1, hw=(40, 40), gaussian_nl=0.0, sp_nl=0.0,
         poisson_noise=True, poisson_alpha=300, gamma1=0.0005, gamma2=1999.0, myLambda=1.0,
         alpha_n=0.9, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="FBS_blur_g", method="A-PnPFBS-DnCNN", deg_op="blur", ch=1, hw=(48, 48), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=1.0, gamma2=0.99, myLambda=1.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=20),
    dict(tag="RED_blur_g", method="A-RED-DnCNN", deg_op="blur", ch=1, hw=(48, 48), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=1.0, gamma2=0.99, myLambda=0.4, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=20),
    dict(tag="B2_blur_g", method="comparisonB-2", deg_op="blur", ch=1, hw=(32, 32), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=1.0, gamma2=0.49, alpha_n=0.95, alpha_s=0.95, r=1.0, m1=4, m2=3, iters=4),
    # comparisonB-4 / comparisonB-5 cannot be recorded: the reference never constructs denoiser_J for
    # them (iteration.py:40 tests for "Proposed"/"DnCNN" in the name) and dies with UnboundLocalError.
    dict(tag="CADMM_blur_g", method="C-PnPADMM-DnCNN", deg_op="blur", ch=1, hw=(32, 32), gaussian_nl=0.0, sp_nl=0.0,
         poisson_noise=True, poisson_alpha=300, gamma1=0.02, gamma2=1.0, myLambda=0.025, m1=10, m2=3,
         gammaInADMMStep1=1.0, alpha_n=0.9, alpha_s=0.95, r=1.0, iters=5),
    dict(tag="CRED_blur_g", method="C-RED-DnCNN", deg_op="blur", ch=1, hw=(32, 32), gaussian_nl=0.0, sp_nl=0.0,
         poisson_noise=True, poisson_alpha=300, gamma1=0.02, gamma2=1.0, myLambda=0.03, m1=8, m2=3,
         gammaInADMMStep1=10.0, alpha_n=0.9, alpha_s=0.95, r=1.0, iters=4),
]

LONG_CASES = [
    dict(tag="LONG_A_blur_g", method="A-Proposed", deg_op="blur", ch=1, hw=(64, 64), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=1200),
    dict(tag="LONG_C_blur_g", method="C-Proposed", deg_op="blur", ch=1, hw=(64, 64), gaussian_nl=0.0, sp_nl=0.0,
         poisson_noise=True, poisson_alpha=100, gamma1=0.0006, gamma2=1 / 0.0006, myLambda=1.0,
         alpha_n=0.9, alpha_s=0.95, r=1.0, iters=1200),
    dict(tag="LONG_B_rs_g", method="B-Proposed", deg_op="random_sampling", ch=1, hw=(64, 64), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=1.0, gamma2=0.49, alpha_n=0.9, alpha_s=0.9, r=0.8, iters=3000),
    dict(tag="LONG_A_blur_c", method="A-Proposed", deg_op="blur", ch=3, hw=(48, 48), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=1200),
]


# TV baselines (iteration.py:88-99,133-140): no denoiser, colour only (operators.py:122-123 hard-code three channels).
# Step sizes of ideas/param_memo.py:21-26 (comparisonA-4: gamma1 = 0.1, gamma2 = 0.99).
TV_CASES = [
    dict(tag="TV_A_blur_c", method="A-PDS-TV", deg_op="blur", ch=3, hw=(32, 48), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.1, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=40),
    dict(tag="TV_A_rs_c", method="A-PDS-TV", deg_op="random_sampling", ch=3, hw=(40, 32), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.1, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=0.8, iters=40),
    dict(tag="TV_FBS_blur_c", method="A-FBS-TV", deg_op="blur", ch=3, hw=(32, 32), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.1, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=40),
    dict(tag="TV_B3_rs_c", method="comparisonB-3", deg_op="random_sampling", ch=3, hw=(32, 32), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=0.1, gamma2=0.49, alpha_n=0.9, alpha_s=0.9, r=0.8, iters=40),
    dict(tag="TV_B3_blur_c", method="comparisonB-3", deg_op="blur", ch=3, hw=(32, 40), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=0.1, gamma2=0.49, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=40),
    dict(tag="LONG_TV_A_blur_c", method="A-PDS-TV", deg_op="blur", ch=3, hw=(48, 48), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.1, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=1200),
]


# Reference-recorded runs at the BASELINE.json shapes and the step sizes of ideas/param_memo.py (:7 cfg1, :44/:50 cfg2,
# :84 cfg3, :62 comparisonB-2).  Stored compactly (gen_base): the final iterate as the float32 the reference returns, the
# observation as float32 (the product ingests fp32; Poisson counts are small integers), x_true regenerated from its seed.
BASE_CASES = [
    dict(tag="BASE_cfg1_A_blur_256", method="A-Proposed", deg_op="blur", ch=1, hw=(256, 256), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=1200),
    dict(tag="BASE_cfg3_C_blur_256", method="C-Proposed", deg_op="blur", ch=1, hw=(256, 256), gaussian_nl=0.0, sp_nl=0.0,
         poisson_noise=True, poisson_alpha=100, gamma1=0.0006, gamma2=1 / 0.0006, myLambda=1.0,
         alpha_n=0.9, alpha_s=0.95, r=1.0, iters=1200),
    dict(tag="BASE_cfg4_A_blur_c1024", method="A-Proposed", deg_op="blur", ch=3, hw=(1024, 1024), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=100),
    dict(tag="BASE_cfg2_B_rs_512", method="B-Proposed", deg_op="random_sampling", ch=1, hw=(512, 512), gaussian_nl=0.01, sp_nl=0.1,
         gamma1=1.0, gamma2=0.49, alpha_n=0.9, alpha_s=0.9, r=0.8, iters=3000),
    dict(tag="BASE_B2_blur_128", method="comparisonB-2", deg_op="blur", ch=1, hw=(128, 128), gaussian_nl=0.01, sp_nl=0.0,
         gamma1=1.0, gamma2=0.49, alpha_n=0.95, alpha_s=0.95, r=1.0, m1=35, m2=5, iters=30),
]

# the "unstable" KAIR DnCNN loops (iteration.py:106-112 colour / 173-180 gray Poisson)
UNSTABLE_CASES = [
    dict(tag="UNS_A_blur_c", method="A-PnPPDS-unstable-DnCNN", arch="dncnn_color_blind", deg_op="blur", ch=3, hw=(40, 48),
         gaussian_nl=0.01, sp_nl=0.0, gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=1.0, iters=30),
    dict(tag="UNS_A_rs_g", method="A-PnPPDS-unstable-DnCNN", arch="dncnn_15", deg_op="random_sampling", ch=1, hw=(48, 48),
         gaussian_nl=0.01, sp_nl=0.0, gamma1=0.99, gamma2=0.99, alpha_n=0.95, alpha_s=0.95, r=0.8, iters=30),
    dict(tag="UNS_C_blur_g", method="C-PnP-unstable-DnCNN", arch="dncnn_15", deg_op="blur", ch=1, hw=(48, 40),
         gaussian_nl=0.0, sp_nl=0.0, poisson_noise=True, poisson_alpha=100, gamma1=0.0006, gamma2=1 / 0.0006, myLambda=1.0,
         alpha_n=0.9, alpha_s=0.95, r=1.0, iters=30),
]


def case_seed(tag):
    return sum(map(ord, tag)) % 997


def gen_base(ref, cases, only_tags=()):
    """One .npz per case (tests/golden/base_<tag>.npz) so a long run can be added without redoing the others."""
    import json
    for case in cases:
        if only_tags and case["tag"] not in only_tags:
            continue
        t = time.time()
        res = run_case(ref, case, snapshots=())
        n = case["iters"]
        x = np.asarray(res[f"x_{n}"])
        assert x.dtype == np.float32, x.dtype                        # the reference's denoiser output (denoiser.py:44)
        obs = np.asarray(res["obs"])
        obs32 = obs.astype(np.float32)
        out = {"x": x, "obs": obs32, "obs_f32_max_abs_err": np.max(np.abs(obs32.astype(np.float64) - obs)),
               "c": res["c"], "psnr": res["psnr"], "seed": case_seed(case["tag"]),
               "x_true_sum": float(np.sum(res["x_true"], dtype=np.float64)), "case": np.array(json.dumps(case)),
               "seconds": time.time() - t}
        s05 = np.asarray(res[f"s05_{n}"])
        if np.any(s05 != 0.5):
            out["s05"] = s05.astype(np.float32)
        save(f"base_{case['tag']}.npz", **out)


def gen_unet(ref):
    """UNet.forward of the reference's own module graph (models/network_unet.py:13-66) for small channel widths.  The class
    cannot be constructed as shipped (network_unet.py:17: load_state_dict(torch.load(file_name)) before any layer exists), so
    exactly that statement is neutralised while __init__ runs; layers, forward and parameter names are the reference's."""
    import torch
    import models.network_unet as nu
    out = {}
    for tag, in_nc, nc, nb, hw in (("g", 1, [8, 16, 24, 32], 2, (24, 40)), ("c", 3, [16, 16, 32, 40], 2, (32, 16)), ("g3", 1, [8, 8, 16, 16], 3, (16, 16))):
        torch.manual_seed(1234 + in_nc + nb)
        orig_load, orig_lsd = torch.load, torch.nn.Module.load_state_dict
        torch.load = lambda *a, **kw: {}
        torch.nn.Module.load_state_dict = lambda self, sd, strict=True: None
        try:
            net = nu.UNet(file_name="", in_nc=in_nc, out_nc=in_nc, nc=nc, nb=nb)
        finally:
            torch.load, torch.nn.Module.load_state_dict = orig_load, orig_lsd
        net.eval()
        with torch.no_grad():
            for prm in net.parameters():                      # default init is tiny through 20 layers: scale up so every layer matters
                prm.mul_(1.5)
            x = torch.rand((2, in_nc) + hw)
            y = net(x)
        for k, v in net.state_dict().items():
            out[f"{tag}/sd/{k}"] = v.numpy()
        out[f"{tag}/x"], out[f"{tag}/y"] = x.numpy(), y.numpy()
        out[f"{tag}/cfg"] = np.array([in_nc, nb] + nc)
        print(f"  unet {tag}: params {sum(p.numel() for p in net.parameters())}, |y - x| max {float((y - x).abs().max()):.3e}")
    save("unet.npz", **out)


def gen_textfile():
    """SUMMARY text-file strings from the reference's own writer (utils/utils_textfile.py) for a fixed `datas` dict."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("ref_textfile", os.path.join(ref_harness.REF_ROOT, "utils", "utils_textfile.py"))
    tf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tf)
    data = {"experimental_settings": {"deg_op": "blur", "gaussian_nl": 0.01, "poisson_alpha": 300, "r": 0.8, "sp_nl": 0.0, "poisson_noise": False},
            "method": {"method": "A-Proposed", "gamma1": 0.99, "gamma2": 0.99, "alpha_n": 0.95, "alpha_s": 0.95, "myLambda": 1, "max_iter": 1200,
                       "m1": 15, "m2": 15, "architecture": "x", "gammaInADMMStep1": 0.1},
            "configs": {"ch": 3, "add_timestamp": True, "result_output": False},
            "summary": {"algorithm": "PnP-PDS", "denoiser": "DnCNN", "Average_PSNR": 31.25, "Average_SSIM": 0.91},
            "results": {0: {"filename": "a.png", "PSNR": 31.0, "SSIM": 0.9, "PSNR_observation": 24.1, "SSIM_observation": 0.6},
                        1: {"filename": "b.png", "PSNR": 31.5, "SSIM": 0.92, "PSNR_observation": 23.9, "SSIM_observation": 0.58}}}
    out = {"data": {k: (v if k != "results" else {str(i): r for i, r in v.items()}) for k, v in data.items()},
           "header": tf.get_csv_header(), "line": tf.get_csv_data(data), "footer": tf.get_csv_footer(data)}
    json.dump(out, open(os.path.join(HERE, "textfile.json"), "w"), indent=1)
    print("wrote textfile.json")


def run_case(ref, case, snapshots):
    """Mirrors main.test_all_images main.py:41-69 (observation synthesis, call into test_iter)."""
    op, un = ref.operators, ref.utils_noise
    ch = case["ch"]
    H, W = case["hw"]
    deg_op, r = case["deg_op"], case.get("r", 1.0)
    img_true = synthetic_image(sum(map(ord, case["tag"])) % 997, ch, H, W)
    phi, adj = op.get_observation_operators(deg_op, KERNEL, r)
    ident, _ = op.get_observation_operators("Id", KERNEL, r)
    nop = phi if deg_op == "random_sampling" else ident
    obs = phi(img_true)
    obs = un.add_gaussian_noise(obs, case["gaussian_nl"], nop)
    if case.get("poisson_noise", False):
        obs = un.apply_poisson_noise(obs, case.get("poisson_alpha", 300))
    obs = un.add_salt_and_pepper_noise(obs, case["sp_nl"], nop)
    x0 = np.copy(obs)
    if case.get("poisson_noise", False):
        x0 = x0 / case.get("poisson_alpha", 300)
    arch = case.get("arch", "DnCNN_nobn_nch_1_nlev_0.01" if ch == 1 else "DnCNN_nobn_nch_3_nlev_0.01")
    path_prox = os.path.join(NN, arch + ".pth")
    res = {}
    # test_iter has no snapshot hook: run it for each snapshot length (cheap at these sizes)
    lens = sorted(set(list(snapshots) + [case["iters"]]))
    lens = [n for n in lens if n <= case["iters"]]
    for n in lens:
        t = time.time()
        x, s05, c, psnr, ssim, avg = ref.iteration.test_iter(
            np.copy(x0), obs, img_true, phi, adj, case["gamma1"], case["gamma2"], case["alpha_s"], case["alpha_n"],
            case.get("myLambda", 1.0), case.get("m1", 15), case.get("m2", 15), case.get("gammaInADMMStep1", 0.1),
            case["gaussian_nl"], case["sp_nl"], case.get("poisson_alpha", 300), path_prox, n, case["method"], ch, r)
        res[f"x_{n}"] = np.asarray(x)
        res[f"s05_{n}"] = np.asarray(s05)
        if n == case["iters"]:
            res["c"], res["psnr"] = c, psnr
        print(f"  {case['tag']} n={n} psnr={psnr[-1]:.4f} c={c[-1]:.3e} ({time.time()-t:.1f}s)")
    res["x_true"], res["obs"], res["x0"] = img_true, np.asarray(obs), np.asarray(x0)
    return res


def gen_loops(ref, cases, fname, snapshots):
    out = {}
    import json
    for case in cases:
        res = run_case(ref, case, snapshots)
        for k, v in res.items():
            out[f"{case['tag']}/{k}"] = v
        out[f"{case['tag']}/case"] = np.array(json.dumps(case))
    save(fname, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true", help="also run the full-iteration-count cases (minutes)")
    ap.add_argument("--only", default="")
    ap.add_argument("--tags", default="", help="with --only base: comma-separated case tags (default: all)")
    a = ap.parse_args()
    ref = ref_harness.load()
    import torch
    torch.set_num_threads(os.cpu_count())
    todo = a.only.split(",") if a.only else ["assets", "ops", "noise", "denoiser", "loops", "tv", "textfile"]
    if "assets" in todo:
        gen_assets(ref)
    if "ops" in todo:
        gen_ops(ref)
    if "noise" in todo:
        gen_noise(ref)
    if "denoiser" in todo:
        gen_denoiser(ref)
    if "loops" in todo:
        gen_loops(ref, LOOP_CASES, "loops.npz", snapshots=(1, 2, 10))
    if "textfile" in todo:
        gen_textfile()
    if "tv" in todo:
        gen_loops(ref, TV_CASES, "tv.npz", snapshots=(1, 2, 10))
    if a.long or "long" in todo:
        gen_loops(ref, LONG_CASES, "long.npz", snapshots=())
    if "unet" in todo:
        gen_unet(ref)
    if "unstable" in todo:
        gen_loops(ref, UNSTABLE_CASES, "unstable.npz", snapshots=(1, 2, 10))
    if "base" in todo:                                               # ~1 h on 8 vCPUs
        gen_base(ref, BASE_CASES, tuple(t for t in a.tags.split(",") if t))


if __name__ == "__main__":
    main()
