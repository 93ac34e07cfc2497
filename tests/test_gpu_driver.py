"""Driver-level drop-in checks on the B200: main.test_all_images / eval_restoration with the reference's dict
API and result layout (main.py:16-123), and the sharded grid search (main.py:125-159) on one rank."""
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2, weights_path
from oracle import pds_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ref_tree(tmp_path, assets):
    """A miniature reference checkout: config dict, nn/, blur_models/, a folder of PNG test images."""
    import cv2
    root = tmp_path / "root"
    (root / "nn").mkdir(parents=True)
    (root / "blur_models").mkdir()
    for arch in ("DnCNN_nobn_nch_1_nlev_0.01", "DnCNN_nobn_nch_3_nlev_0.01"):
        shutil.copy(weights_path(arch), root / "nn" / (arch + ".pdsw"))
    np.save(root / "blur_models" / "blur_1.npy", assets["blur_1"])
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    for b in range(3):
        rgb = (O.synthetic_image(10 + b, 3, 48, 64) * 255).round().astype(np.uint8)
        cv2.imwrite(str(imgs / f"im{b}.png"), np.moveaxis(rgb, 0, -1)[..., ::-1])
    res = tmp_path / "out"
    res.mkdir()
    return dict(path_test=str(imgs) + os.sep, path_result=str(res) + os.sep, pattern_red="*.png", root_folder=str(root) + os.sep)


def test_test_all_images_matches_per_image_oracle(ref_tree, assets):
    from pnp_pds_b200 import main as pmain
    settings = {"gaussian_nl": 0.01, "sp_nl": 0, "poisson_noise": False, "deg_op": "blur", "r": 0.8}
    method = {"method": "ours-A", "architecture": "DnCNN_nobn_nch_3_nlev_0.01", "max_iter": 8, "gamma1": 0.99, "gamma2": 0.99, "alpha_n": 0.95}
    datas = pmain.test_all_images(settings, method, {"ch": 3}, config=ref_tree)
    assert set(datas) == {"experimental_settings", "method", "configs", "results", "summary"}
    assert datas["method"]["m1"] == 15 and datas["experimental_settings"]["poisson_alpha"] == 300          # defaults filled in
    assert set(datas["summary"]) == {"Average_PSNR", "PSNR", "Average_SSIM", "SSIM", "Average_time", "Cpu_time", "algorithm", "denoiser"}
    assert datas["summary"]["algorithm"] == "PnP-PDS" and datas["summary"]["denoiser"] == "DnCNN"
    assert len(datas["results"]) == 3
    r0 = datas["results"][0]
    assert set(r0) == {"filename", "c_evolution", "PSNR_evolution", "SSIM_evolution", "GROUND_TRUTH", "OBSERVATION", "RESULT",
                       "REMOVED_SPARSE", "PSNR", "SSIM", "CPU_time", "PSNR_observation", "SSIM_observation"}
    assert r0["filename"] == "im0.png" and r0["RESULT"].shape == (3, 48, 64) and r0["c_evolution"].shape == (8,)
    # one image against the oracle, with the observation the driver synthesised (reference order main.py:49-64)
    from pnp_pds_b200.models.weights import load_weights
    w = load_weights(weights_path("DnCNN_nobn_nch_3_nlev_0.01"))
    img = r0["GROUND_TRUTH"]
    x0, obs = O.synthesize_observation(img, "blur", assets["blur_1"], 0.8, 0.01, 0, False, 300)
    assert np.max(np.abs(obs - r0["OBSERVATION"])) < 5e-6                      # fp32 GPU blur vs float64 stencil; same noise draw
    phi, adj = O.make_operators("blur", assets["blur_1"], 0.8)
    den = lambda z: O.dncnn_forward(w.layers, z, w.slope, w.residual_sign, w.clamp)
    xr, _, c, psnr, _ = O.pds_iterations(r0["OBSERVATION"], r0["OBSERVATION"], img, phi, adj, den, 0.99, 0.99, 1, 0.95, 1, 0.01, 0, 300, 8,
                                         "A-Proposed", 0.8)
    assert rel_l2(r0["RESULT"], xr) < 1e-4
    assert abs(r0["PSNR"] - psnr[-1]) < 0.01
    assert abs(datas["summary"]["Average_PSNR"] - np.mean([datas["results"][i]["PSNR"] for i in range(3)])) < 1e-12
    assert any(f.startswith("DATA_") for f in os.listdir(ref_tree["path_result"]))


def test_eval_restoration_legacy_signature(ref_tree):
    from pnp_pds_b200 import main as pmain
    psnr = pmain.eval_restoration(gaussian_nl=0.01, sp_nl=0.0, poisson_noise=False, poisson_alpha=0, max_iter=5, gamma1=0.99, gamma2=0.99,
                                  r=1, alpha_n=0.95, alpha_s=0.95, myLambda=1, result_output=False,
                                  architecture="DnCNN_nobn_nch_3_nlev_0.01", deg_op="blur", method="ours-A", ch=3, config=ref_tree)
    assert 20 < psnr < 40


def test_grid_search_single_rank_matches_individual_runs(assets):
    from pnp_pds_b200 import iteration, main as pmain, operators
    images = [O.synthetic_image(b, 1, 40, 40) for b in range(3)]
    grid = [dict(alpha_n=0.82 + 0.06 * i, gamma1=0.99, gamma2=0.99) for i in range(3)] + [dict(alpha_n=0.9, gamma1=0.5, gamma2=1.5)]
    settings = {"gaussian_nl": 0.02, "sp_nl": 0, "poisson_noise": False, "deg_op": "blur", "r": 0.8}
    common = {"method": "A-Proposed", "max_iter": 6}
    path = weights_path("DnCNN_nobn_nch_1_nlev_0.01")
    table = pmain.grid_search(images, grid, settings, common, 1, assets["blur_1"], path, batch_size=5)   # 12 items in batches of 5
    assert table.shape == (3, 4, 3)
    phi, adj = operators.get_observation_operators("blur", assets["blur_1"], 0.8)
    x0, obs = pmain.synthesize_observation(images[1], phi, "blur", 0.02, 0, False, 300)
    p = dict(gamma1=0.5, gamma2=1.5, alpha_s=1, alpha_n=0.9, myLambda=1, gaussian_nl=0.02, sp_nl=0, poisson_alpha=300, r=0.8)
    one = iteration.run_batch(x0[None], obs[None], images[1][None], phi, adj, p, path, 6, "A-Proposed", 1)
    assert abs(table[1, 3, 0] - one["psnr"][-1, 0]) < 1e-9 and abs(table[1, 3, 2] - one["c"][-1, 0]) < 1e-12
    assert len({round(v, 6) for v in table[0, :, 0]}) == 4                                         # grid points differ


def test_sweep_driver_writes_summary_file(ref_tree):
    """main.main (main.py:125-159) over a two-experiment list: SUMMARY(<timestamp>).txt with the reference's layout, one line per
    experiment; includes a TV baseline (no denoiser) next to ours-A."""
    from pnp_pds_b200 import main as pmain
    from pnp_pds_b200.utils import utils_textfile as tf
    settings = {"gaussian_nl": 0.01, "sp_nl": 0, "poisson_noise": False, "deg_op": "blur", "r": 0.8}
    ex = [{"settings": settings, "configs": {"ch": 3},
           "method": {"method": "A-Proposed", "architecture": "DnCNN_nobn_nch_3_nlev_0.01", "max_iter": 3, "gamma1": 0.99, "gamma2": 0.99, "alpha_n": 0.9}},
          {"settings": settings, "configs": {"ch": 3},
           "method": {"method": "A-PDS-TV", "architecture": "DnCNN_nobn_nch_3_nlev_0.01", "max_iter": 3, "gamma1": 0.125, "gamma2": 0.99, "alpha_n": 0.9}}]
    path = pmain.main(config=ref_tree, experiments=ex)
    lines = open(path).read().split("\n")
    assert os.path.basename(path).startswith("SUMMARY(") and lines[0] + "\n" == tf.get_csv_header()
    assert len(lines) == 5 and lines[4] == ""                      # header, 2 experiments, footer, trailing newline
    a, t = lines[1].split(","), lines[2].split(",")
    assert a[0] == "blur" and a[3] == "A-Proposed" and a[4] == "PnP-PDS" and a[5] == "DnCNN"
    assert t[3] == "A-PDS-TV" and t[4] == "PDS" and t[5] == ""
    assert 10 < float(a[6]) < 50 and 10 < float(t[6]) < 50        # average PSNR of three images
    assert len(a) == 17 + 4 * 3 + 1                                 # 17 columns, 4 values per image, trailing comma
    assert lines[3] == "im0.png,im1.png,im2.png,"
