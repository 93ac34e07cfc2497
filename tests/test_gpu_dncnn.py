"""DnCNN forward parity on the B200: both conv engines vs the reference's outputs (golden) and the oracle."""
import numpy as np
import pytest

from conftest import weights_path
from oracle import pds_oracle as O

pytestmark = pytest.mark.gpu

SIMPLE = ["DnCNN_nobn_nch_1_nlev_0.01", "DnCNN_nobn_nch_3_nlev_0.01", "DnCNN_nobn_nch_1_nlev_0.009"]
KAIR = [("dncnn_15", 1, 17), ("dncnn_color_blind", 3, 20), ("dncnn3", 1, 20)]
# fp32 accumulation; activations carried between the 20 layers as fp16 hi+lo (2^-22, SIMT engine) or as fp16 plus an
# e4m3 first-order correction (~2^-16 operand precision, tcgen05 engine; tools/emulate_split.py predicts 2.5e-6..4.4e-6
# for the simple_CNN checkpoints and 0.9e-5..3.5e-5 for the KAIR ones, against 4.5e-5..1.1e-3 for one plain fp16 pass)
TOL = 1e-5
KAIR_TOL = {"simt": 5e-5, "tcgen05": 1e-4}   # no clamps + ReLU: activations an order of magnitude larger than simple_CNN's


@pytest.mark.parametrize("engine", ["simt", "tcgen05"])
@pytest.mark.parametrize("arch", SIMPLE)
def test_simple_cnn_vs_reference(g_den, arch, engine):
    from pnp_pds_b200.models.denoiser import Denoiser
    ch = 3 if "nch_3" in arch else 1
    den = Denoiser(weights_path(arch), ch, conv_engine=engine)
    x = g_den[f"{arch}_x"]
    y = den.denoise(x)
    ref = g_den[f"{arch}_y"]
    assert y.dtype == np.float32 and y.shape == ref.shape
    err = float(np.max(np.abs(y - ref)))
    print(f"{arch} {engine}: max abs err {err:.3e}")
    assert err < TOL
    assert y.min() >= 0 and y.max() <= 1


@pytest.mark.parametrize("engine", ["simt", "tcgen05"])
@pytest.mark.parametrize("arch,ch,nb", KAIR)
def test_kair_dncnn_vs_reference(g_den, arch, ch, nb, engine):
    import torch
    from pnp_pds_b200.models.denoiser import Denoiser
    from pnp_pds_b200.models.weights import load_weights
    den = Denoiser(load_weights(weights_path(arch)), ch, conv_engine=engine)
    x = g_den[f"{arch}_x"]
    y = den.denoise_batch(x[None])[0]
    err = float(np.max(np.abs(y - g_den[f"{arch}_y"])))
    print(f"{arch} {engine}: max abs err {err:.3e}")
    assert err < KAIR_TOL[engine]


def test_kair_class_interface(g_den):
    import torch
    from pnp_pds_b200.models.network_dncnn import DnCNN
    net = DnCNN(in_nc=1, out_nc=1, nc=64, nb=17, act_mode="R", model_path=weights_path("dncnn_15"))
    x = torch.from_numpy(g_den["dncnn_15_x"])          # unbatched gray (1,H,W), as iteration.py:108 passes it
    y = net(x)
    assert tuple(y.shape) == tuple(x.shape)
    assert float((y - torch.from_numpy(g_den["dncnn_15_y"])).abs().max()) < 2 * TOL
    with pytest.raises(RuntimeError):
        DnCNN(in_nc=1, out_nc=1, nc=64, nb=20, act_mode="R", model_path=weights_path("dncnn_15"))


@pytest.mark.parametrize("shape", [(1, 50, 37), (3, 33, 70), (1, 16, 8), (1, 7, 5)])
def test_engines_agree_ragged_shapes(shape):
    """Tile edges: H not a multiple of 16, W not a multiple of 8, image smaller than a tile."""
    from pnp_pds_b200.models.denoiser import Denoiser
    from pnp_pds_b200.models.weights import load_weights
    C, H, W = shape
    arch = SIMPLE[1] if C == 3 else SIMPLE[0]
    w = load_weights(weights_path(arch))
    x = np.random.default_rng(5).random(shape).astype(np.float32)
    ref = O.dncnn_forward(w.layers, x, w.slope, w.residual_sign, w.clamp)
    for engine in ("simt", "tcgen05"):
        y = Denoiser(w, C, conv_engine=engine).denoise(x if C == 3 else x[0])
        err = float(np.max(np.abs(y.reshape(shape) - ref)))
        print(shape, engine, err)
        assert err < TOL, (shape, engine)


def test_batch_matches_single_and_chunking():
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    w = load_weights(weights_path(SIMPLE[0]))
    rng = np.random.default_rng(9)
    x = rng.random((5, 1, 48, 40)).astype(np.float32)
    outs = []
    for chunk in (0, 2):                                # default chunk (all 5 at once) and 2 images per pass
        with Engine(5, 1, 48, 40, conv_engine="tcgen05", denoiser_chunk=chunk) as e:
            e.load_dncnn(w)
            outs.append(e.dncnn_forward(e.to_device(x)).cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    with Engine(1, 1, 48, 40, conv_engine="tcgen05") as e:
        e.load_dncnn(w)
        for b in range(5):
            one = e.dncnn_forward(e.to_device(x[b])).cpu().numpy()
            assert np.array_equal(one[0], outs[0][b])


def test_full_size_linearity_free_checks():
    """256x256 (config 1 size): engines agree with each other; output stays in [0,1]."""
    from pnp_pds_b200.models.denoiser import Denoiser
    x = O.synthetic_image(0, 1, 256, 256) + 0.02 * np.random.default_rng(1).standard_normal((256, 256)).astype(np.float32)
    a = Denoiser(weights_path(SIMPLE[0]), 1, conv_engine="simt").denoise(x)
    b = Denoiser(weights_path(SIMPLE[0]), 1, conv_engine="tcgen05").denoise(x)
    assert float(np.max(np.abs(a - b))) < TOL
    assert a.min() >= 0 and a.max() <= 1


@pytest.mark.parametrize("shape", [(12, 1, 256, 256), (24, 3, 120, 128), (16, 1, 72, 384), (1, 1, 23, 128), (2, 3, 9, 256),
                                   (2, 1, 40, 321), (3, 3, 19, 200), (1, 1, 64, 130)])
def test_row_streaming_body_kernel_matches_tile_kernels(shape):
    """Large launches whose width splits into 128-pixel strips run the row-streaming body kernel (dncnn_roll.cu):
    same operands and the same MMA order per output element, so the results are bit-identical.  Covers partial last bands,
    a lone strip whose pair partner lies outside the image (W = 128), an odd strip count (W = 384) and widths that are not
    a multiple of the strip width (the last strip hangs over the right edge: TMA zero-fill in, masked stores out)."""
    from pnp_pds_b200 import _lib
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    B, C, H, W = shape
    lib = _lib.load()
    assert lib.pds_debug_roll_band_rows(B, H, W, 1) > 0, "narrower than one strip"
    w = load_weights(weights_path(SIMPLE[1] if C == 3 else SIMPLE[0]))
    x = np.random.default_rng(11).random(shape).astype(np.float32)
    outs = {}
    # force / forbid row streaming (the cost model may pick either); bit 8: the row-streaming kernel reads e4m3(a) from HBM
    # instead of rebuilding it from the fp16 plane in shared memory (then every layer stores it)
    for name, variant in (("roll", 64), ("roll_hbm", 64 | 256), ("tile", 128)):
        with Engine(B, C, H, W, conv_engine="tcgen05") as e:
            e.load_dncnn(w)
            e.set_tc_variant(variant)
            outs[name] = e.dncnn_forward(e.to_device(x)).cpu().numpy()
    assert np.array_equal(outs["roll"], outs["roll_hbm"])      # same operands bit for bit, same MMA order
    err = float(np.max(np.abs(outs["roll"] - outs["tile"])))
    print(shape, "roll vs tile", err)
    assert err < 2e-6                      # fp32 accumulation order only
    for b in (0, B - 1):
        ref = O.dncnn_forward(w.layers, x[b], w.slope, w.residual_sign, w.clamp)
        assert float(np.max(np.abs(outs["roll"][b] - ref))) < TOL


@pytest.mark.parametrize("shape", [(1, 1, 48, 24), (3, 3, 50, 37), (2, 1, 130, 70)])
def test_tile_kernels_one_and_two_cta_agree(shape):
    """The 1-CTA and the CTA-pair tile kernels issue the same MMAs in the same order: bit-identical outputs, including an
    odd tile count (the pair's second tile is a dead duplicate) and ragged edges."""
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    B, C, H, W = shape
    w = load_weights(weights_path(SIMPLE[1] if C == 3 else SIMPLE[0]))
    x = np.random.default_rng(21).random(shape).astype(np.float32)
    outs = []
    for variant in (16 | 512, 32 | 512):           # force the 1-CTA / the 2-CTA tile kernel, one launch per body layer
        with Engine(B, C, H, W, conv_engine="tcgen05") as e:
            e.load_dncnn(w)
            e.set_tc_variant(variant)
            outs.append(e.dncnn_forward(e.to_device(x)).cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    ref = O.dncnn_forward(w.layers, x[0], w.slope, w.residual_sign, w.clamp)
    assert float(np.max(np.abs(outs[0][0] - ref))) < TOL


@pytest.mark.parametrize("shape,arch", [((1, 1, 256, 256), SIMPLE[0]), ((1, 1, 48, 24), SIMPLE[0]), ((3, 3, 50, 37), SIMPLE[1]),
                                        ((2, 1, 130, 70), SIMPLE[0]), ((1, 1, 16, 8), SIMPLE[0]), ((5, 1, 96, 200), "dncnn_15"),
                                        ((1, 3, 200, 120), "dncnn_color_blind"), ((1, 1, 512, 512), SIMPLE[0])])
def test_chain_kernel_matches_per_layer_tile_kernels(shape, arch):
    """All body layers in one persistent launch with tile-level dataflow between layers (dncnn_chain.cu) against one
    conv_tc2_kernel launch per layer: the same MMAs in the same order per output element -> bit-identical, for a single
    256^2 / 512^2 image, fewer units than clusters (16 x 8: one tile), odd tile counts, ragged edges, several images, the
    17-layer and the colour KAIR networks; and again on repeated calls (the per-unit flags count up across launches)."""
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    B, C, H, W = shape
    w = load_weights(weights_path(arch))
    rng = np.random.default_rng(31)
    xs = [rng.random(shape).astype(np.float32) for _ in range(3)]
    with Engine(B, C, H, W, conv_engine="tcgen05") as e:
        e.load_dncnn(w)
        e.set_tc_variant(128 | 512)                 # tile kernels, one launch per layer
        ref = [e.dncnn_forward(e.to_device(x)).cpu().numpy() for x in xs]
        n0 = e.kernel_launches
        e.dncnn_forward(e.to_device(xs[0]))
        per_layer = e.kernel_launches - n0
        e.set_tc_variant(128)                       # tile kernels' territory -> the chain kernel
        n0 = e.kernel_launches
        out = [e.dncnn_forward(e.to_device(x)).cpu().numpy() for x in xs]
        chained = (e.kernel_launches - n0) // 3
    assert per_layer == w.depth and chained == 3, (per_layer, chained)      # first + ONE body launch + last
    for a, b in zip(ref, out):
        assert np.array_equal(a, b)
    o = O.dncnn_forward(w.layers, xs[0][0] if C == 3 else xs[0][0, 0], w.slope, w.residual_sign, w.clamp)
    assert float(np.max(np.abs(out[0][0] - o.reshape(out[0][0].shape)))) < (1e-4 if arch.startswith("dncnn") else TOL)


def test_chain_kernel_partial_last_chunk_and_back_to_back_calls():
    """A batch that does not divide into denoiser chunks: the last chunk has fewer tile pairs than the others, so its launch
    starts from zeroed flags (launch_conv_chain); 20 back-to-back passes stay bit-identical to the per-layer kernels."""
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    B, C, H, W = 5, 1, 64, 72
    w = load_weights(weights_path(SIMPLE[0]))
    x = np.random.default_rng(33).random((B, C, H, W)).astype(np.float32)
    with Engine(B, C, H, W, conv_engine="tcgen05", denoiser_chunk=2) as e:
        e.load_dncnn(w)
        e.set_tc_variant(128 | 512)
        ref = e.dncnn_forward(e.to_device(x)).cpu().numpy()
        e.set_tc_variant(128)
        xd = e.to_device(x)
        for _ in range(20):
            out = e.dncnn_forward(xd)
        assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("shape,variant,name", [((2, 1, 40, 256), 64, "roll_d"), ((1, 1, 24, 256), 64 | 256, "roll_hbm"), ((3, 1, 50, 37), 128 | 512, "tc2"),
                                                ((3, 1, 50, 37), 128, "chain"), ((1, 3, 64, 64), 128, "chain_colour")])
def test_body_kernels_are_deterministic_run_to_run(shape, variant, name):
    """Race canary (compute-sanitizer is closed on the GPU pool, profiles/r02_sanitizer.txt): 12 forward passes of the same input
    through each body-kernel family give bit-identical outputs."""
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    B, C, H, W = shape
    w = load_weights(weights_path(SIMPLE[1] if C == 3 else SIMPLE[0]))
    x = np.random.default_rng(41).random(shape).astype(np.float32)
    with Engine(B, C, H, W, conv_engine="tcgen05") as e:
        e.load_dncnn(w)
        e.set_tc_variant(variant)
        xd = e.to_device(x)
        first = e.dncnn_forward(xd).clone()
        for _ in range(11):
            assert bool((e.dncnn_forward(xd) == first).all()), name


@pytest.mark.parametrize("shape,arch", [((2, 1, 64, 64), SIMPLE[0]), ((3, 3, 50, 36), SIMPLE[1]), ((1, 1, 23, 128), SIMPLE[0]),
                                        ((2, 3, 9, 260), SIMPLE[1]), ((1, 1, 40, 132), SIMPLE[0]), ((2, 1, 7, 4), SIMPLE[0]),
                                        ((1, 3, 33, 321), SIMPLE[1]), ((2, 3, 40, 48), "dncnn_color_blind"), ((1, 1, 48, 48), "dncnn_15")])
def test_first_layer_kernels_agree(shape, arch):
    """The tap-shifted first-layer kernel (one K = 16 MMA per tap over 32-byte pixel records, input windows by TMA, bias folded into
    the centre tap; the default whenever W % 4 == 0) against the im2col kernel (tc_variant bit 15) and the oracle: strips hanging
    over the right edge (W = 36, 132, 260), images narrower than a strip, one-row bands, W % 4 != 0 (both variants then run the
    im2col kernel: bit-identical), the KAIR networks (no input clamp, ReLU)."""
    from pnp_pds_b200.engine import Engine
    from pnp_pds_b200.models.weights import load_weights
    B, C, H, W = shape
    w = load_weights(weights_path(arch))
    x = (np.random.default_rng(31).random(shape).astype(np.float32) * 1.2 - 0.1)     # exercises the input clamp where there is one
    outs = []
    for variant in (0, 32768):
        with Engine(B, C, H, W, conv_engine="tcgen05") as e:
            e.load_dncnn(w)
            e.set_tc_variant(variant)
            outs.append(e.dncnn_forward(e.to_device(x)).cpu().numpy())
    err = float(np.max(np.abs(outs[0] - outs[1])))
    print(shape, arch, "tap-shifted vs im2col", err)
    if W % 4:
        assert np.array_equal(outs[0], outs[1])
    tol = 2e-4 if "dncnn" in arch else TOL      # KAIR networks: no clamps, activations an order of magnitude larger
    assert err < tol                    # three fp16 products in both; accumulation order and the bias split differ
    for b in (0, B - 1):
        ref = O.dncnn_forward(w.layers, x[b], w.slope, w.residual_sign, w.clamp)
        e_new, e_old = float(np.max(np.abs(outs[0][b] - ref))), float(np.max(np.abs(outs[1][b] - ref)))
        print("   vs oracle: tap-shifted", e_new, "im2col", e_old)
        assert e_new < tol and e_old < tol
