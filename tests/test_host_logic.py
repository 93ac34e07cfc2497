"""CPU-only checks of the host side: C-ABI surface, weight conversion, bit-exact masks/noise of the
product code against the reference's recorded outputs, argument handling, sharding (gloo, world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, weights_path


def test_library_loads_and_exports_every_declared_symbol():
    from pnp_pds_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "pnp_pds.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pds_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pnp_pds.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert lib.pds_abi_version() == 1
    assert isinstance(lib.pds_device_count(), int)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pnp_pds_b200._lib import PdsError
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200 import operators
    with pytest.raises(PdsError):
        Engine(1, 1, 8, 8)
    phi, _ = operators.get_observation_operators("random_sampling", None, 0.8)
    with pytest.raises(PdsError):
        phi(np.ones((8, 8)))


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "pnp-pds_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "pds_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_weight_blob_roundtrip_and_shapes():
    from pnp_pds_b200.models.weights import DnCNNWeights, load_weights
    for arch, (depth, c, slope, rs, clamp) in {
        "DnCNN_nobn_nch_1_nlev_0.01": (20, 1, 0.01, 1.0, True), "DnCNN_nobn_nch_3_nlev_0.01": (20, 3, 0.01, 1.0, True),
        "DnCNN_nobn_nch_1_nlev_0.009": (20, 1, 0.01, 1.0, True), "dncnn_15": (17, 1, 0.0, -1.0, False),
        "dncnn3": (20, 1, 0.0, -1.0, False), "dncnn_color_blind": (20, 3, 0.0, -1.0, False)}.items():
        w = load_weights(weights_path(arch))
        assert (w.depth, w.c_in, w.c_out, w.n_ch, w.residual_sign, w.clamp) == (depth, c, c, 64, rs, clamp)
        assert abs(w.slope - slope) < 1e-7
        n_par = sum(a.size + b.size for a, b in w.layers)
        if depth == 20:
            assert n_par == (665921 if c == 1 else 668227)
        w2 = DnCNNWeights.from_blob(w.to_blob())
        assert all(np.array_equal(a, c_) and np.array_equal(b, d) for (a, b), (c_, d) in zip(w.layers, w2.layers))
    with pytest.raises(ValueError):
        DnCNNWeights.from_blob(b"nope" + bytes(60))


@pytest.mark.skipif(not os.path.isdir("/root/reference/nn"), reason="reference checkpoints not mounted")
def test_pth_conversion_matches_committed_blobs_and_blocks_foreign_globals(tmp_path):
    from pnp_pds_b200.models.weights import load_pth
    for arch in ("DnCNN_nobn_nch_1_nlev_0.01", "dncnn_15"):
        w = load_pth(f"/root/reference/nn/{arch}.pth")
        assert w.to_blob() == open(weights_path(arch), "rb").read()
    import pickle, torch

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    p = tmp_path / "evil.pth"
    torch.save(Evil(), str(p), _use_new_zipfile_serialization=False)
    with pytest.raises(pickle.UnpicklingError):
        load_pth(str(p))


def test_legacy_checkpoint_first_pickle_goes_through_the_allow_list(tmp_path):
    """torch's legacy loader reads the magic number / protocol / sys-info / storage keys with pickle_module.load: a file
    whose FIRST pickle carries a REDUCE on os.system must be rejected, not executed."""
    import pickle
    from pnp_pds_b200.models.weights import load_pth
    marker = tmp_path / "executed"

    class Evil:
        def __reduce__(self):
            return (os.system, (f"touch {marker}",))
    p = tmp_path / "evil_first.pth"
    with open(p, "wb") as f:
        pickle.dump(Evil(), f, protocol=2)          # where the magic number belongs
        pickle.dump(1001, f, protocol=2)
        pickle.dump({}, f, protocol=2)
    with pytest.raises(pickle.UnpicklingError):
        load_pth(str(p))
    assert not marker.exists()


@pytest.mark.parametrize("H,W,r", [(32, 32, 0.8), (64, 64, 0.5), (48, 20, 0.7), (256, 256, 0.8), (512, 512, 0.8), (1024, 1024, 0.8)])
def test_product_mask_bit_exact(g_ops, H, W, r):
    from pnp_pds_b200.operators import sampling_mask
    ref = np.unpackbits(g_ops[f"mask_{H}_{W}_{r}"])[: H * W].reshape(H, W)
    assert np.array_equal(sampling_mask(H, W, r), ref)


def test_product_noise_bit_exact(g_noise):
    from pnp_pds_b200.operators import sampling_mask
    from pnp_pds_b200.utils import utils_noise as un
    ident = lambda z: z
    rs = lambda z: z * sampling_mask(z.shape[-2], z.shape[-1], 0.8)
    img, imgc = g_noise["img_g"], g_noise["img_c"]
    assert np.array_equal(un.add_gaussian_noise(img, 0.01, ident), g_noise["gauss_g_id"])
    assert np.array_equal(un.add_gaussian_noise(rs(img), 0.01, rs), g_noise["gauss_g_rs"])
    assert np.array_equal(un.apply_poisson_noise(img, 100), g_noise["poisson_g"])
    assert un.apply_poisson_noise(img, 100).dtype == np.int64
    assert np.array_equal(un.add_salt_and_pepper_noise(img, 0.1, ident), g_noise["sp_g_id"])
    assert np.array_equal(un.add_salt_and_pepper_noise(rs(img), 0.1, rs), g_noise["sp_g_rs"])
    assert np.array_equal(un.add_salt_and_pepper_noise(rs(imgc), 0.1, rs), g_noise["sp_c_rs"])
    assert np.array_equal(un.add_salt_and_pepper_noise(img, 0.0, ident), g_noise["sp_g_zero"])


def test_arg_parsing_defaults_and_aliases():
    from pnp_pds_b200.engine import canonical_method, RESIDENT_METHODS
    from pnp_pds_b200.utils.utils_method_master import get_algorithm_denoiser
    from pnp_pds_b200.utils.utils_parse_args import parse_args_configs, parse_args_exp, parse_args_method
    from pnp_pds_b200.utils.utils_unparse_args import unparse_args_method
    assert parse_args_exp({}) == (0, 0, False, 300, "blur", 0.8)                               # utils_parse_args.py:5-10
    assert parse_args_method({}) == ("ours-A", "DnCNN_nobn_nch_3_nlev_0.01", 10, 1, 1, 1, 1, 1, 15, 15, 0.1)
    assert parse_args_configs({}) == (3, True, False)
    assert parse_args_method({"gamma1": 0.5, "m2": 3})[3] == 0.5
    assert unparse_args_method(*parse_args_method({}))["gammaInADMMStep1"] == 0.1
    for old, new in (("ours-A", "A-Proposed"), ("ours-B", "B-Proposed"), ("ours-C", "C-Proposed"),
                     ("comparisonA-1", "A-PnPFBS-DnCNN"), ("comparisonA-6", "A-RED-DnCNN")):
        assert canonical_method(old) == new and new in RESIDENT_METHODS
    assert get_algorithm_denoiser("ours-B") == ("PnP-PDS", "DnCNN")
    assert get_algorithm_denoiser("A-RED-DnCNN") == ("RED-SD", "DnCNN")
    assert get_algorithm_denoiser("whatever") == ("unknown algorithm", "unknown denoiser")


def test_item_params_quirks():
    from pnp_pds_b200.iteration import item_params
    n = 512 * 512
    # B-Proposed passes r to both projections (iteration.py:56,58); cfg2 numbers from SURVEY §8 a-8/a-9
    pb = item_params("B", n, 1.0, 0.49, 0.9, 0.9, 1, 0.01, 0.1, 300, 0.8)
    assert abs(pb["epsilon"] - 3.497) < 1e-3 and abs(pb["eta"] - 9437.18) < 0.01
    # A-Proposed never passes r (iteration.py:52): epsilon uses r = 1 even for random_sampling
    pa = item_params("A", n, 0.99, 0.99, 0.9, 0.9, 1, 0.01, 0.0, 300, 0.8)
    assert abs(pa["epsilon"] - np.sqrt(n) * 0.9 * 0.01) < 1e-9


def test_metrics_from_traces():
    from pnp_pds_b200.engine import metrics_from_traces
    tr = np.zeros((2, 1, 5))
    tr[:, 0, 1] = [4.0, 1.0]
    tr[:, 0, 2] = [16.0, 16.0]
    tr[:, 0, 3] = [0.01 * 100, 0.0001 * 100]
    c, psnr = metrics_from_traces(tr, 100)
    assert np.allclose(c[:, 0], [0.5, 0.25]) and np.allclose(psnr[:, 0], [20.0, 40.0])
    from pnp_pds_b200.engine import ssim_from_traces
    tr[1, 0, 4] = 0.5 * 3 * (20 - 6) * (10 - 6)
    s3 = ssim_from_traces(tr, 3, 20, 10)
    assert np.isnan(s3[0, 0]) and abs(s3[1, 0] - 0.5) < 1e-15
    tr[1, 0, 4] = 0.25 * 20 * (10 - 6)
    assert abs(ssim_from_traces(tr, 1, 20, 10)[1, 0] - 0.25) < 1e-15      # gray: H rows x (W-6) one-dimensional windows


def test_ssim_restatement_matches_oracle_and_basic_properties():
    from oracle import pds_oracle as O
    from pnp_pds_b200.utils.utils_eval import eval_psnr, eval_ssim
    rng = np.random.default_rng(0)
    a = rng.random((3, 40, 32))
    b = np.clip(a + 0.05 * rng.standard_normal(a.shape), 0, 1)
    assert abs(eval_ssim(a, b) - O.eval_ssim(a, b)) < 1e-12
    assert abs(eval_ssim(a, a) - 1.0) < 1e-12
    g = rng.random((40, 32))
    assert 0 < eval_ssim(g, np.clip(g + 0.05 * rng.standard_normal(g.shape), 0, 1)) < 1
    assert abs(eval_psnr(a, b) - O.eval_psnr(a, b)) < 1e-12


def test_ssim_restatement_against_scipy_uniform_filter_formulation():
    """scikit-image is absent (SSIM parity stays formally unpinned), but the primitive its structural_similarity is built on,
    scipy.ndimage.uniform_filter, is here.  This spells out skimage 0.22's algorithm with it — per channel along axis 0
    (channel_axis=0: a gray (H, W) image is H one-dimensional signals), size-7 uniform filter in 'reflect' mode, sample
    covariance NP/(NP-1), K1=.01, K2=.03, crop (win-1)//2, float64 mean — and checks the product's valid-window form."""
    from scipy.ndimage import uniform_filter
    from pnp_pds_b200.utils.utils_eval import eval_ssim

    def skimage_like(im1, im2):
        data_range = im2.max() - im2.min()
        vals = []
        for ch in range(im1.shape[0]):
            a, b = im1[ch].astype(np.float64), im2[ch].astype(np.float64)
            NP = 7 ** a.ndim
            cov = NP / (NP - 1)
            ux, uy = uniform_filter(a, size=7), uniform_filter(b, size=7)
            uxx, uyy, uxy = uniform_filter(a * a, size=7), uniform_filter(b * b, size=7), uniform_filter(a * b, size=7)
            vx, vy, vxy = cov * (uxx - ux * ux), cov * (uyy - uy * uy), cov * (uxy - ux * uy)
            C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
            S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
            S = S[tuple(slice(3, -3) for _ in range(a.ndim))]
            vals.append(S.mean(dtype=np.float64))
        return float(np.mean(vals))

    rng = np.random.default_rng(5)
    for shape in ((3, 40, 32), (24, 33), (3, 7, 9)):
        a = rng.random(shape)
        b = np.clip(a + 0.1 * rng.standard_normal(shape), 0, 1)
        assert abs(eval_ssim(a, b) - skimage_like(a, b)) < 1e-10, shape


def test_shard_range_partitions():
    from pnp_pds_b200.parallel import shard_range
    for n in (0, 1, 7, 256, 2560):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
from pnp_pds_b200.parallel import init_distributed, shard_range, gather_rows
rank, local_rank, world = init_distributed("gloo")
n_items = 7
lo, hi = shard_range(n_items, rank, world)
rows = np.stack([[i, 10.0 * i + 0.5, -i] for i in range(lo, hi)]) if hi > lo else np.zeros((0, 3))
out = gather_rows(rows, n_items)
exp = np.stack([[i, 10.0 * i + 0.5, -i] for i in range(n_items)])
assert out.shape == (n_items, 3) and np.array_equal(out, exp), out
print("rank", rank, "ok")
"""


def test_gather_rows_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    import socket
    last = None
    for attempt in range(3):                             # the rendezvous of a freshly started pair occasionally times out on a loaded host
        with socket.socket() as sk:                      # a free port (a fixed one may linger in TIME_WAIT between runs)
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        last = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                               "127.0.0.1", "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=240)
        if last.returncode == 0 and "rank 0 ok" in last.stdout and "rank 1 ok" in last.stdout:
            return
    assert False, last.stdout + last.stderr


def test_row_streaming_band_selection_cost_model():
    """pds_debug_roll_band_rows is host arithmetic (148 SMs assumed without a device): rows per CTA pair when the
    row-streaming kernels serve the launch (large launches at least one 128-pixel strip wide), 0 for the tile kernels."""
    from pnp_pds_b200 import _lib
    lib = _lib.load()
    f = lib.pds_debug_roll_band_rows
    assert f(8, 1024, 1024, 0) == 443                     # cfg4: 8 images per denoiser pass, 32768 strip-pair rows over 74 pairs
    assert f(1, 1024, 1024, 0) == 56                      # one 1024x1024 image still fills the CTA pairs
    assert f(1, 256, 256, 0) == 0                         # a single small image: halo rows dominate, tiles win
    assert f(1, 512, 512, 0) == 14                        # 1024 strip-pair rows over 74 pairs
    assert f(1, 100, 100, 0) == 0 and f(1, 100, 100, 1) == 0      # narrower than one strip: never
    assert f(16, 321, 481, 0) > 0                         # ragged widths: 481 = 3.76 strips, still cheaper than tiles
    assert f(16, 200, 130, 0) == 0                        # ... but not when most of a strip pair hangs over the edge
    assert f(1, 256, 256, 1) >= 8                         # forced (tests)


def test_converter_lane_mapping_is_complete_and_conflict_free():
    """Host restatement of the lane -> (pixel, 16-channel quarter) mapping of the converter warps in conv_roll_d_kernel
    (dncnn_roll.cu): every 16-byte chunk of the 130-pixel row box is produced exactly once, and within a quarter-warp (the
    unit in which 16-byte shared-memory accesses are served) the two fp16 reads and the two e4m3 writes of SWIZZLE_128B rows
    land in eight distinct 16-byte columns, i.e. all 32 banks once."""
    row_pix, full = 130, 16
    seen = set()
    for k in range(full + 1):
        for lane in range(32):
            if k == full and lane >= 4 * (row_pix - 8 * full):
                continue
            pig = ((lane >> 3) ^ 5) if ((lane >> 2) & 1) else (lane >> 3)
            p = k * 8 + pig if k < full else 8 * full + (lane >> 2)
            qd = lane & 3
            assert 0 <= p < row_pix and (p, qd) not in seen
            seen.add((p, qd))
        if k == full:
            break
        for q in range(4):                                 # quarter-warps of task k
            cols = {"ld0": [], "ld1": [], "st_a8": [], "st_lo": []}
            for lane in range(8 * q, 8 * q + 8):
                pig = ((lane >> 3) ^ 5) if ((lane >> 2) & 1) else (lane >> 3)
                p, qd = k * 8 + pig, lane & 3
                sw = p & 7
                cols["ld0"].append((2 * qd) ^ sw)
                cols["ld1"].append((2 * qd + 1) ^ sw)
                cols["st_a8"].append(qd ^ sw)
                cols["st_lo"].append((4 + qd) ^ sw)
            for name, c in cols.items():
                assert sorted(c) == list(range(8)), (k, q, name, c)
    assert len(seen) == 4 * row_pix


def test_band_walk_covers_every_row_once():
    """Host restatement of BandWalk + launch_conv_mid_roll (dncnn_roll.cu): the CTA pairs' contiguous shares, cut into bands at
    the column boundaries, cover every (image, strip pair, row) exactly once and differ by at most one share in length."""
    from pnp_pds_b200 import _lib
    f = _lib.load().pds_debug_roll_band_rows
    for nimg, H, W in ((8, 1024, 1024), (1, 512, 512), (12, 256, 256), (16, 321, 481), (1, 23, 128), (2, 9, 256)):
        rpc = f(nimg, H, W, 1)
        npx = (W + 255) // 256
        total = nimg * npx * H
        assert rpc == max(8, -(-total // 74))
        nclusters = -(-total // rpc)
        assert nclusters <= 74
        seen = np.zeros((nimg, npx, H), dtype=np.int32)
        steps = []
        for cid in range(nclusters):
            g, end, n = cid * rpc, min(total, (cid + 1) * rpc), 0
            while g < end:
                col, yb = divmod(g, H)
                rb = min(H - yb, end - g)
                img, px = divmod(col, npx)
                seen[img, px, yb:yb + rb] += 1
                g += rb
                n += rb + 2
            steps.append(n)
        assert (seen == 1).all()
        assert max(steps) <= rpc + 2 * (2 + rpc // H)


def test_summary_textfile_matches_reference_format(tmp_path):
    """utils_textfile against strings produced by the reference's writer (tests/golden/textfile.json, recorded here)."""
    import json
    from pnp_pds_b200.utils import utils_textfile as tf
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "textfile.json")))
    data = dict(g["data"])
    data["results"] = {int(k): v for k, v in data["results"].items()}
    assert tf.get_csv_header() == g["header"]
    assert tf.get_csv_data(data) == g["line"]
    assert tf.get_csv_footer(data) == g["footer"]
    path = tmp_path / "SUMMARY.txt"
    tf.touch_textfile(path)
    tf.write_textfile(path, data)
    tf.add_footer_textfile(path, data)
    assert path.read_text() == g["header"] + g["line"] + "\n" + g["footer"] + "\n"


def test_sweep_experiment_list():
    """main.sweep_experiments mirrors main.py:130-153: 5 x 2 x (2 x 10 + 2 x 10) experiments."""
    from pnp_pds_b200.main import sweep_experiments
    ex = sweep_experiments()
    assert len(ex) == 5 * 2 * 40
    assert ex[0]["method"] == {"method": "A-Proposed", "max_iter": 1200, "gamma1": 0.99, "gamma2": 0.99, "alpha_n": 0.8 + 0.02}
    tv = [e for e in ex if e["method"]["method"] == "A-PDS-TV"][0]
    assert tv["method"]["gamma1"] == 0.125
    lams = [e["method"]["myLambda"] for e in ex if e["method"]["method"] == "A-RED-DnCNN"][:10]
    assert lams[-1] == 1.99 and abs(lams[0] - 0.2) < 1e-12
    assert [e["method"]["max_iter"] for e in ex if e["settings"]["deg_op"] == "random_sampling"][0] == 3000
