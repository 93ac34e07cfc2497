"""KAIR UNet forward (models/network_unet.py:13-66): the oracle against outputs of the reference's own module
graph (tests/golden/unet.npz, make_golden.py gen_unet), and the CUDA forward (csrc/unet.cu through pds_unet_*) against both."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2

TAGS = ["g", "c", "g3"]


def _case(g, tag):
    in_nc, nb, *nc = [int(v) for v in g[f"{tag}/cfg"]]
    pre = f"{tag}/sd/"
    sd = {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
    return in_nc, nb, nc, sd, g[f"{tag}/x"], g[f"{tag}/y"]


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_unet_matches_reference_module_graph(tag):
    from oracle.unet_oracle import unet_forward
    from pnp_pds_b200.models.network_unet import layer_keys, layer_shapes
    in_nc, nb, nc, sd, x, y = _case(load_golden("unet.npz"), tag)
    # the reference's parameter names and shapes are the ones the product's blob writer expects
    assert sorted(sd) == sorted(k + s for k, _ in layer_keys(nb) for s in (".weight", ".bias"))
    for (k, _), shp in zip(layer_keys(nb), layer_shapes(in_nc, in_nc, nc, nb)):
        assert tuple(sd[k + ".weight"].shape) == tuple(shp), k
    out = unet_forward(sd, x, nb)
    assert out.shape == y.shape and np.max(np.abs(out - y)) < 2e-5       # fp32 reference vs float64 restatement
    one = unet_forward(sd, x[0], nb)
    assert np.max(np.abs(one - y[0])) < 2e-5


def test_unet_blob_layout_and_argument_checks():
    from pnp_pds_b200.models.network_unet import UNet, layer_shapes, random_state_dict, to_blob
    sd = random_state_dict(1, 1, (8, 16, 24, 32), 2, seed=3)
    blob = to_blob(sd, 1, 1, (8, 16, 24, 32), 2)
    n = sum(int(np.prod(s)) for s in layer_shapes(1, 1, [8, 16, 24, 32], 2))
    nbias = sum(v.size for k, v in sd.items() if k.endswith(".bias"))
    assert blob[:4] == b"PDSU" and len(blob) == 48 + 4 * (n + nbias)
    with pytest.raises(ValueError):
        to_blob({**sd, "m_tail.weight": sd["m_tail.weight"][:, :4]}, 1, 1, (8, 16, 24, 32), 2)
    with pytest.raises(NotImplementedError):
        UNet(state_dict=sd, nc=(8, 16, 24, 32), act_mode="L")
    with pytest.raises(ValueError):
        UNet(state_dict=sd, in_nc=1, out_nc=3, nc=(8, 16, 24, 32))
    with pytest.raises(ValueError):
        UNet(nc=(8, 16, 24, 32))                                          # no weights: the reference ships none


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_unet_forward_matches_reference(tag):
    from pnp_pds_b200.models.network_unet import UNet
    in_nc, nb, nc, sd, x, y = _case(load_golden("unet.npz"), tag)
    net = UNet(in_nc=in_nc, out_nc=in_nc, nc=nc, nb=nb, state_dict=sd)
    out = net(x)
    e = float(np.max(np.abs(out - y)))
    print(f"unet {tag}: max abs err vs reference {e:.2e}, rel_l2 {rel_l2(out, y):.2e}")
    assert out.dtype == np.float32 and out.shape == y.shape and e < 2e-5
    assert np.array_equal(net(x[1]), out[1])                              # unbatched (C,H,W) input, second handle
    import torch
    t = net(torch.from_numpy(x))
    assert isinstance(t, torch.Tensor) and np.array_equal(t.numpy(), out)
    net.close()


@pytest.mark.gpu
def test_unet_reference_widths_against_oracle():
    """The reference's default widths (64, 128, 256, 512), nb = 2, colour, 2 x 3 x 64 x 96 — random parameters (there is no trained
    UNet), against the float64 oracle; plus the C-ABI's argument checks."""
    import ctypes as C
    from oracle.unet_oracle import unet_forward
    from pnp_pds_b200 import _lib
    from pnp_pds_b200.models.network_unet import UNet, random_state_dict
    sd = random_state_dict(3, 3, (64, 128, 256, 512), 2, seed=5)
    x = np.random.default_rng(6).random((2, 3, 64, 96)).astype(np.float32)
    net = UNet(in_nc=3, out_nc=3, state_dict=sd)
    out = net(x)
    ref = unet_forward(sd, x, 2)
    e = rel_l2(out, ref)
    print(f"unet 64/128/256/512: rel_l2 {e:.2e}, max abs {np.max(np.abs(out - ref)):.2e}, |y-x| max {np.max(np.abs(ref - x)):.2e}")
    assert e < 1e-5
    with pytest.raises(_lib.PdsError):
        net(np.zeros((3, 60, 96), np.float32))                            # H not divisible by 8
    net.close()
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.PdsUnetConfig(1, 1, 1, (C.c_int32 * 4)(8, 8, 8, 8), 2, 16, 16, 0)
    assert lib.pds_unet_create(C.byref(cfg), C.byref(h)) == 0
    x1 = __import__("torch").zeros((1, 1, 16, 16), device="cuda")
    assert lib.pds_unet_forward(h, C.c_void_p(x1.data_ptr()), C.c_void_p(x1.data_ptr() + 4), None) != 0
    assert "not loaded" in lib.pds_last_error().decode()
    assert lib.pds_unet_load(h, b"XXXX" + bytes(60), 64) != 0 and "PDSU" in lib.pds_last_error().decode()
    assert lib.pds_unet_destroy(h) == 0
