"""Error behaviour of the C ABI (include/pnp_pds.h): every misuse returns non-zero with a message in pds_last_error and
leaves the handle usable; nothing throws across the boundary, nothing falls back to the CPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import weights_path

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from pnp_pds_b200._lib import PdsConfig
    d = dict(batch=1, channels=1, height=32, width=32, method=0, deg_op=0, max_iter=4, reserved=0, device=0, denoiser_chunk=0)
    d.update(kw)
    return PdsConfig(*[d[n] for n, _ in PdsConfig._fields_])


def _err(lib):
    return lib.pds_last_error().decode()


@pytest.mark.parametrize("kw,needle", [
    (dict(batch=0), "bad shape"), (dict(channels=2), "channels"), (dict(method=99), "unknown method"),
    (dict(deg_op=7), "unknown deg_op"), (dict(max_iter=0), "max_iter"), (dict(device=99), "device"),
    (dict(method=8, channels=1), "colour"),                       # TV baselines need three channels
    (dict(reserved=1), "reserved"),
    (dict(channels=3, height=40000, width=40000), "2^31"),
])
def test_create_rejects_bad_configs(kw, needle):
    from pnp_pds_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.pds_create(C.byref(_cfg(**kw)), C.byref(h)) != 0
    assert needle in _err(lib), _err(lib)
    assert not h.value


def test_call_order_and_argument_checks():
    import torch
    from pnp_pds_b200 import _lib
    from pnp_pds_b200._lib import PdsItemParams
    from pnp_pds_b200.models.weights import load_weights
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.pds_create(C.byref(_cfg(deg_op=1)), C.byref(h)) == 0
    try:
        x = torch.full((1, 1, 32, 32), 0.5, device="cuda")
        out = torch.empty_like(x)
        p = lambda t: C.c_void_p(t.data_ptr())
        # nothing configured yet
        assert lib.pds_run(h, 1, None) != 0 and "pds_set_problem" in _err(lib)
        assert lib.pds_phi(h, p(x), p(out), None) != 0 and "blur kernel not set" in _err(lib)
        assert lib.pds_dncnn_forward(h, p(x), p(out), None) != 0 and "weights not loaded" in _err(lib)
        # bad blur kernels
        k = np.zeros((4, 4))
        assert lib.pds_set_blur_kernel(h, k.ctypes.data_as(C.POINTER(C.c_double)), 4) != 0 and "odd" in _err(lib)
        k = np.zeros((3, 3))
        assert lib.pds_set_blur_kernel(h, k.ctypes.data_as(C.POINTER(C.c_double)), 3) != 0 and "all zeros" in _err(lib)
        k[1, 1] = 1.0
        assert lib.pds_set_blur_kernel(h, k.ctypes.data_as(C.POINTER(C.c_double)), 3) == 0
        assert lib.pds_phi(h, p(x), p(x), None) != 0 and "in place" in _err(lib)
        assert lib.pds_phi(h, p(x), p(out), None) == 0 and torch.allclose(out, x)          # delta kernel = identity
        k2 = np.zeros((5, 5)); k2[2, 3] = 1.0                                                # replacing the kernel: a one-pixel shift
        assert lib.pds_set_blur_kernel(h, k2.ctypes.data_as(C.POINTER(C.c_double)), 5) == 0
        xr = torch.rand((1, 1, 32, 32), device="cuda")
        assert lib.pds_phi(h, p(xr), p(out), None) == 0 and torch.equal(out, torch.roll(xr, 1, dims=3))
        assert lib.pds_set_blur_kernel(h, k.ctypes.data_as(C.POINTER(C.c_double)), 3) == 0
        # weights: wrong magic, truncated blob, wrong channel count
        assert lib.pds_load_dncnn(h, b"XXXX" + bytes(60), 64) != 0 and "PDSW" in _err(lib)
        blob = load_weights(weights_path("DnCNN_nobn_nch_1_nlev_0.01")).to_blob()
        assert lib.pds_load_dncnn(h, blob, len(blob) - 4) != 0 and "size mismatch" in _err(lib)
        blob3 = load_weights(weights_path("DnCNN_nobn_nch_3_nlev_0.01")).to_blob()
        assert lib.pds_load_dncnn(h, blob3, len(blob3)) != 0 and "channel count" in _err(lib)
        assert lib.pds_load_dncnn(h, blob, len(blob)) == 0
        assert lib.pds_load_dncnn(h, blob, len(blob)) != 0 and "already loaded" in _err(lib)
        # parameters and iteration budget
        prm = (PdsItemParams * 2)()
        assert lib.pds_set_item_params(h, prm, 2) != 0 and "n must be 1 or batch" in _err(lib)
        prm[0] = PdsItemParams(0.99, 0.99, 0.3, 0.0, 1.0, 300.0)
        assert lib.pds_set_item_params(h, prm, 1) == 0
        assert lib.pds_set_problem(h, p(x), p(x), None, None) == 0
        assert lib.pds_run(h, 5, None) != 0 and "max_iter" in _err(lib)
        assert lib.pds_run(h, 4, None) == 0 and lib.pds_iterations_done(h) == 4
        assert lib.pds_run(h, 1, None) != 0                                               # budget used up
        tr = np.zeros((4, 1, 5))
        assert lib.pds_get_traces(h, tr.ctypes.data_as(C.c_void_p), 3, None) != 0 and "too small" in _err(lib)
        assert lib.pds_get_traces(h, tr.ctypes.data_as(C.c_void_p), tr.size, None) == 0 and np.all(np.isfinite(tr))
    finally:
        assert lib.pds_destroy(h) == 0
    assert lib.pds_run(None, 1, None) != 0 and "null handle" in _err(lib)
    h2 = C.c_void_p()
    assert lib.pds_create(C.byref(_cfg(deg_op=0)), C.byref(h2)) == 0
    try:
        k = np.zeros((3, 3)); k[1, 1] = 1.0
        assert lib.pds_set_blur_kernel(h2, k.ctypes.data_as(C.POINTER(C.c_double)), 3) != 0 and "deg_op" in _err(lib)
        assert lib.pds_debug_set_conv_engine(h2, 7) != 0 and "unknown conv engine" in _err(lib)
    finally:
        assert lib.pds_destroy(h2) == 0
