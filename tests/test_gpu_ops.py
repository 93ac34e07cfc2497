"""Stand-alone operator parity on the B200 (through the C ABI) against the golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import pds_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from pnp_pds_b200 import operators
    return operators


@pytest.mark.parametrize("tag", ["g32", "c24", "g48x20"])
def test_blur_vs_reference(ops, assets, g_ops, tag):
    phi, adj = ops.get_observation_operators("blur", assets["blur_1"], 0.8)
    x = g_ops[f"blur_{tag}_x"]
    assert np.max(np.abs(phi(x) - g_ops[f"blur_{tag}_phi"])) < 2e-6      # fp32 stencil vs float64 FFT, O(1) values
    assert np.max(np.abs(adj(x) - g_ops[f"blur_{tag}_adj"])) < 2e-6


@pytest.mark.parametrize("shape", [(1, 256, 256), (3, 100, 72), (1, 512, 512)])
def test_blur_properties_full_size(ops, assets, shape):
    from pnp_pds_b200.engine import Engine
    import torch
    C, H, W = shape
    with Engine(2, C, H, W, deg_op="blur") as e:
        e.set_blur_kernel(assets["blur_1"])
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn((2, C, H, W), device="cuda", generator=g)
        y = torch.randn((2, C, H, W), device="cuda", generator=g)
        lhs = (e.phi(x).double() * y.double()).sum().item()
        rhs = (x.double() * e.phi_adj(y).double()).sum().item()
        assert abs(lhs - rhs) < 1e-5 * (x.numel() ** 0.5) * 10       # <Phi x, y> = <x, Phi^T y>
        ones = torch.ones((2, C, H, W), device="cuda")
        assert (e.phi(ones) - 1).abs().max().item() < 1e-5           # kernel sums to 1
        # linearity
        z = e.phi(2 * x - 3 * y) - (2 * e.phi(x) - 3 * e.phi(y))
        assert z.abs().max().item() < 1e-4
    # against the oracle on one plane
    xs = x[0, 0].cpu().numpy()
    ref = O.blur_phi(xs.astype(np.float64), assets["blur_1"])
    phi, _ = ops.get_observation_operators("blur", assets["blur_1"], 1)
    assert np.max(np.abs(phi(xs) - ref)) < 1e-5


@pytest.mark.parametrize("H,W,r", [(32, 32, 0.8), (48, 20, 0.7), (512, 512, 0.8)])
def test_random_sampling(ops, g_ops, H, W, r):
    ref_mask = np.unpackbits(g_ops[f"mask_{H}_{W}_{r}"])[: H * W].reshape(H, W)
    assert np.array_equal(ops.sampling_mask(H, W, r), ref_mask)        # bit-exact integer work
    phi, adj = ops.get_observation_operators("random_sampling", None, r)
    x = np.random.default_rng(0).random((3, H, W)).astype(np.float32)
    y = phi(x)
    assert np.array_equal(y, (x * ref_mask).astype(np.float64))
    assert np.array_equal(adj(y.astype(np.float32)), y)                 # idempotent, self-adjoint
    xg = x[0]
    assert np.array_equal(phi(xg), xg * ref_mask) and phi(xg).dtype == xg.dtype


def test_random_sampling_golden(ops, g_ops):
    phi, _ = ops.get_observation_operators("random_sampling", None, 0.8)
    assert np.max(np.abs(phi(g_ops["rs_c16_x"]) - g_ops["rs_c16_out"])) < 1e-7
    assert np.array_equal(phi(g_ops["rs_g16_x"]), g_ops["rs_g16_out"])


def test_identity(ops):
    phi, adj = ops.get_observation_operators("Id", None, 0.8)
    x = np.ones((4, 4))
    assert phi(x) is x and adj(x) is x


def test_proj_l2_ball(ops, g_ops):
    a_n, g_nl, sp, r = g_ops["l2_params"]
    out = ops.proj_l2_ball(g_ops["l2_x"], a_n, g_nl, sp, g_ops["l2_b"], r)
    assert rel_l2(out, g_ops["l2_out"]) < 1e-6
    x_in = g_ops["l2_b"] + 1e-4 * g_ops["l2_x"]
    out = ops.proj_l2_ball(x_in, a_n, g_nl, sp, g_ops["l2_b"], r)
    assert np.max(np.abs(out - g_ops["l2_out_inside"])) < 1e-7
    # idempotent, lands on the sphere
    eps = O.l2_ball_radius(g_ops["l2_x"].size, a_n, g_nl, sp, r)
    p1 = ops.proj_l2_ball(g_ops["l2_x"], a_n, g_nl, sp, g_ops["l2_b"], r)
    assert abs(np.linalg.norm(p1 - g_ops["l2_b"]) - eps) < 1e-5 * max(eps, 1)


def test_proj_l1_ball(ops, g_ops):
    a_s, sp, r = g_ops["l1_params"]
    for key in ("l1", "l1b"):
        out = ops.proj_l1_ball(g_ops[f"{key}_x"], a_s, sp, r)
        assert np.max(np.abs(out - g_ops[f"{key}_out"])) < 2e-6, key
        eta = O.l1_ball_radius(out.size, a_s, sp, r)
        assert abs(np.abs(out).sum() - eta) < 1e-4 * eta
    x_small = g_ops["l1_x"] * 1e-3
    assert np.max(np.abs(ops.proj_l1_ball(x_small, a_s, sp, r) - g_ops["l1_out_inside"])) < 1e-9    # inside: identity


@pytest.mark.parametrize("n,scale,sp", [(512 * 512, 0.3, 0.1), (3 * 1024 * 1024, 1.0, 0.05), (1000, 2.0, 0.2), (4096, 0.0, 0.1)])
def test_proj_l1_ball_large(ops, n, scale, sp):
    rng = np.random.default_rng(n)
    z = (rng.standard_normal(n) * scale).astype(np.float32)
    z[::97] *= 8
    out = ops.proj_l1_ball(z, 0.9, sp, 0.8)
    ref = O.proj_l1_ball(z.astype(np.float64), 0.9, sp, 0.8)
    assert np.max(np.abs(out - ref)) < 5e-6 * max(1.0, float(np.abs(z).max()))
    eta = O.l1_ball_radius(n, 0.9, sp, 0.8)
    if np.abs(z).sum() > eta:
        assert abs(np.abs(out).sum() - eta) < 2e-4 * eta


def test_proj_l1_eta_zero(ops):
    z = np.random.default_rng(3).standard_normal(5000).astype(np.float32)
    assert np.all(ops.proj_l1_ball(z, 0.9, 0.0, 1.0) == 0)              # sp_nl = 0 -> eta = 0 -> s == 0 (SURVEY a-18)


def test_prox_gkl(ops, g_ops):
    gamma, alpha = g_ops["gkl_params"]
    out = ops.prox_GKL(g_ops["gkl_x"], gamma, alpha, g_ops["gkl_x0"])
    assert np.max(np.abs(out - g_ops["gkl_out"]) / np.maximum(1.0, np.abs(g_ops["gkl_out"]))) < 1e-6
    # very negative argument: cancellation-free branch stays accurate
    x = np.full((8, 8), -5000.0)
    x0 = np.full((8, 8), 3.0)
    ref = O.prox_gkl(x, 0.01, 100.0, x0)
    got = ops.prox_GKL(x, 0.01, 100.0, x0)
    assert np.max(np.abs(got - ref) / ref) < 1e-5


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (1, 3, 96, 200), (40, 1, 128, 256)])
def test_blur1_compile_time_tap_list_matches_generic_stencil(assets, shape):
    """blur_models/blur_1.mat has its own stencil instantiation (non-zero pattern known at compile time, pds_blur.cu SPEC = 1 / 2);
    same taps in the same order as the generic kernels (tc_variant bit 13) -> bit-identical Phi, Phi^T and loop iterates, for the
    32 x 32 and the 128 x 32 tile variants.  A kernel with a different pattern (one tap removed) takes the generic path."""
    from pnp_pds_b200.engine import Engine
    B, C, H, W = shape
    x = np.random.default_rng(5).random(shape).astype(np.float32)
    outs = {}
    for name, variant in (("spec", 0), ("generic", 8192)):
        with Engine(B, C, H, W, method="A", deg_op="blur", max_iter=4) as e:
            e.set_blur_kernel(assets["blur_1"])
            e.set_tc_variant(variant)
            xd = e.to_device(x)
            outs[name] = (e.phi(xd).cpu().numpy(), e.phi_adj(xd).cpu().numpy())
    assert np.array_equal(outs["spec"][0], outs["generic"][0]) and np.array_equal(outs["spec"][1], outs["generic"][1])
    h2 = np.array(assets["blur_1"], dtype=np.float64)
    a, b = np.argwhere(h2 != 0)[3]
    h2[a, b] = 0.0
    with Engine(B, C, H, W, method="A", deg_op="blur", max_iter=4) as e:
        e.set_blur_kernel(h2)
        xd = e.to_device(x)
        y = e.phi(xd).cpu().numpy()
    ref = np.stack([O.blur_phi(x[i].astype(np.float64), h2) for i in range(min(B, 2))])
    assert np.max(np.abs(y[: min(B, 2)] - ref.reshape(y[: min(B, 2)].shape))) < 2e-6
