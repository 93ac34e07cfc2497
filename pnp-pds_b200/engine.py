"""Host-side owner of one libpnp_pds handle: a batch of B independent restorations resident on one B200.

PyTorch is used only for device memory and streams; all compute goes through the C ABI.
Reference: the state that iteration.test_iter sets up per call (iteration.py:23-41) lives here for
the whole batch; one `run()` = the `for i in range(max_iter)` loop (iteration.py:44-189).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import CONV_ENGINES, DEG_OPS, METHODS, TRACE_WIDTH, PdsConfig, PdsError, PdsItemParams
from .models.weights import DnCNNWeights

# method name -> engine method id (both vocabularies)
METHOD_ALIASES = {
    "ours-A": "A-Proposed", "ours-B": "B-Proposed", "ours-C": "C-Proposed",
    "comparisonA-1": "A-PnPFBS-DnCNN", "comparisonA-6": "A-RED-DnCNN",
    "comparisonA-7": "A-PnPPDS-unstable-DnCNN", "comparisonC-4": "C-PnP-unstable-DnCNN",
    "comparisonC-2": "C-PnPADMM-DnCNN", "comparisonC-3": "C-RED-DnCNN",
    "comparisonA-2": "A-PnPPDS-BM3D", "comparisonC-1": "C-PnPPDS-BM3D",
    "comparisonA-4": "A-PDS-TV",                       # ideas/param_memo.py:21-26 (gamma1 = 0.1: the TV primal-dual baseline)
}
RESIDENT_METHODS = {
    "A-Proposed": "A", "B-Proposed": "B", "C-Proposed": "C",
    "A-PnPFBS-DnCNN": "FBS", "A-RED-DnCNN": "RED",
    # ADMM cross-checks (algorithm/admm.py)
    "comparisonB-2": "ADMM_B2", "C-PnPADMM-DnCNN": "ADMM_C", "C-RED-DnCNN": "RED_C",
    # the "unstable" KAIR variants run the same PDS loop with a different denoiser epilogue
    "A-PnPPDS-unstable-DnCNN": "A", "C-PnP-unstable-DnCNN": "C",
    # TV baselines: no denoiser, colour images only (iteration.py:88-99,133-140)
    "A-PDS-TV": "TV_A", "A-FBS-TV": "TV_FBS", "comparisonB-3": "TV_B3",
}
DENOISER_FREE = {"TV_A", "TV_FBS", "TV_B3"}


def canonical_method(method: str) -> str:
    return METHOD_ALIASES.get(method, method)


def l2_ball_radius(n: int, alpha_n: float, gaussian_nl: float, sp_nl: float, r: float = 1.0) -> float:
    """operators.py:104"""
    return float(np.sqrt(n * (1 - sp_nl)) * r * alpha_n * gaussian_nl)


def l1_ball_radius(n: int, alpha_s: float, sp_nl: float, r: float = 1.0) -> float:
    """operators.py:96"""
    return float(alpha_s * n * sp_nl * r * 0.5)


def _torch():
    import torch
    return torch


def _require_cuda(device: int):
    torch = _torch()
    if not torch.cuda.is_available():
        raise PdsError("no CUDA device visible: pnp_pds_b200 has no CPU fallback")
    return torch.device("cuda", device)


class Engine:
    def __init__(self, batch: int, channels: int, height: int, width: int, method: str = "A", deg_op: str = "Id",
                 max_iter: int = 1, conv_engine: str = "tcgen05", device: int | None = None, denoiser_chunk: int = 0):
        self.lib = _lib.load()
        torch = _torch()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = _require_cuda(device)
        self.B, self.C, self.H, self.W = int(batch), int(channels), int(height), int(width)
        self.n = self.C * self.H * self.W
        self.method, self.deg_op, self.max_iter = method, deg_op, int(max_iter)
        cfg = PdsConfig(self.B, self.C, self.H, self.W, METHODS[method], DEG_OPS[deg_op], self.max_iter,
                        0, self.device.index, int(denoiser_chunk))
        h = C.c_void_p()
        _lib.check(self.lib.pds_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.conv_engine = conv_engine
        if CONV_ENGINES[conv_engine] != 0:      # the fp32 CUDA-core cross-check engine: a test hook, not a configuration field
            _lib.check(self.lib.pds_debug_set_conv_engine(self._h, CONV_ENGINES[conv_engine]))
        self.shape = (self.B, self.C, self.H, self.W)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.pds_destroy(self._h)
            self._h = None
        self._staging = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ set-up
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def set_blur_kernel(self, h: np.ndarray):
        h = np.ascontiguousarray(h, dtype=np.float64)
        if h.ndim != 2 or h.shape[0] != h.shape[1]:
            raise ValueError("blur kernel must be square")
        _lib.check(self.lib.pds_set_blur_kernel(self._h, h.ctypes.data_as(C.POINTER(C.c_double)), h.shape[0]))

    def set_mask(self, mask: np.ndarray):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        if m.shape != (self.H, self.W):
            raise ValueError("mask shape must be (H, W)")
        _lib.check(self.lib.pds_set_mask(self._h, m.ctypes.data_as(C.POINTER(C.c_uint8))))

    def set_params(self, params: Sequence[dict] | dict):
        """params: one dict (broadcast) or B dicts with gamma1,gamma2,epsilon,eta,lam,alpha."""
        if isinstance(params, dict):
            params = [params]
        if len(params) not in (1, self.B):
            raise ValueError("need 1 or B parameter sets")
        arr = (PdsItemParams * len(params))()
        for i, p in enumerate(params):
            arr[i] = PdsItemParams(float(p.get("gamma1", 1.0)), float(p.get("gamma2", 1.0)), float(p.get("epsilon", 0.0)),
                                   float(p.get("eta", 0.0)), float(p.get("lam", 1.0)), float(p.get("alpha", 1.0)))
        _lib.check(self.lib.pds_set_item_params(self._h, arr, len(params)))

    def set_ssim(self, mode: str | int):
        """'none' | 'all' (every iteration, as the reference) | 'final' (last iteration of each run)."""
        mode = {"none": 0, "all": 1, "final": 2}.get(mode, mode)
        _lib.check(self.lib.pds_set_ssim(self._h, int(mode)))

    def set_admm(self, m1: int, m2: int, gamma_step1: float):
        """m1, m2, gammaInADMMStep1 of iteration.test_iter (inner trip counts / step of algorithm/admm.py)."""
        _lib.check(self.lib.pds_set_admm(self._h, int(m1), int(m2), float(gamma_step1)))

    def load_dncnn(self, weights: DnCNNWeights | bytes):
        blob = weights.to_blob() if isinstance(weights, DnCNNWeights) else bytes(weights)
        buf = C.create_string_buffer(blob, len(blob))
        _lib.check(self.lib.pds_load_dncnn(self._h, C.cast(buf, C.c_void_p), len(blob)))

    BODY_KERNELS = {0: "conv_mid_simt_kernel (fp32 CUDA-core cross-check engine)",
                    1: "roll::conv_roll_d_kernel (row-streaming cta_group::2 body layer, e4m3(a) operand rebuilt on chip from the fp16 row)",
                    2: "two::conv_tc2_kernel (cta_group::2 tile kernel, one launch per layer)",
                    3: "conv_tc_kernel<64> (1-CTA tile kernel, one launch per layer)",
                    4: "chain::conv_chain_kernel (all body layers in one persistent cta_group::2 launch, tile-level dataflow between layers)"}

    def body_kernel(self, nimg: int = 0) -> int:
        """Which kernel serves the 64->64 body layers (pds_debug_body_kernel): key of BODY_KERNELS."""
        return int(self.lib.pds_debug_body_kernel(self._h, int(nimg)))

    def set_tc_variant(self, v: int):
        _lib.check(self.lib.pds_debug_set_tc_variant(self._h, int(v)))

    # ------------------------------------------------------------------ helpers
    def to_device(self, a) -> "torch.Tensor":
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32)).to(self.device)
        t = t.reshape(self.shape).contiguous()
        return t

    def empty(self):
        torch = _torch()
        return torch.empty(self.shape, dtype=torch.float32, device=self.device)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    # ------------------------------------------------------------------ stand-alone operators (device tensors)
    def phi(self, x):
        out = self.empty()
        _lib.check(self.lib.pds_phi(self._h, self._p(x), self._p(out), self._stream()))
        return out

    def phi_adj(self, x):
        out = self.empty()
        _lib.check(self.lib.pds_phi_adj(self._h, self._p(x), self._p(out), self._stream()))
        return out

    def proj_l2_ball(self, x, center, epsilon: float):
        out = self.empty()
        _lib.check(self.lib.pds_proj_l2_ball(self._h, self._p(x), self._p(center), float(epsilon), self._p(out), self._stream()))
        return out

    def proj_l1_ball(self, x, eta: float):
        out = self.empty()
        _lib.check(self.lib.pds_proj_l1_ball(self._h, self._p(x), float(eta), self._p(out), self._stream()))
        return out

    def prox_gkl(self, x, x0, gamma: float, alpha: float):
        out = self.empty()
        _lib.check(self.lib.pds_prox_gkl(self._h, self._p(x), self._p(x0), float(gamma), float(alpha), self._p(out), self._stream()))
        return out

    def dncnn_forward(self, x):
        out = self.empty()
        _lib.check(self.lib.pds_dncnn_forward(self._h, self._p(x), self._p(out), self._stream()))
        return out

    # ------------------------------------------------------------------ resident loop
    def set_problem(self, x0, obs, x_true=None):
        self._keep = (self.to_device(x0), self.to_device(obs), None if x_true is None else self.to_device(x_true))
        a, b, c = self._keep
        _lib.check(self.lib.pds_set_problem(self._h, self._p(a), self._p(b), self._p(c), self._stream()))

    def run(self, n_iter: int):
        _lib.check(self.lib.pds_run(self._h, int(n_iter), self._stream()))

    @property
    def iterations_done(self) -> int:
        return int(self.lib.pds_iterations_done(self._h))

    def state(self, want_s=True, want_y=False):
        x = self.empty()
        s = self.empty() if want_s else None
        y = self.empty() if want_y else None
        _lib.check(self.lib.pds_get_state(self._h, self._p(x), self._p(s), self._p(y), self._stream()))
        return x, s, y

    def traces(self) -> np.ndarray:
        """float64 array [iterations_done, B, TRACE_WIDTH = 5] of (||t||^2, ||dx||^2, ||x||^2, ||x+ - x_true||^2, SSIM map sum);
        see PDS_TRACE_WIDTH in include/pnp_pds.h."""
        it = self.iterations_done
        out = np.zeros((it, self.B, TRACE_WIDTH), dtype=np.float64)
        if it:
            _lib.check(self.lib.pds_get_traces(self._h, out.ctypes.data_as(C.c_void_p), out.size, self._stream()))
        return out

    # ------------------------------------------------------------------ page-locked staging of the host-buffer entry point
    STAGING_LIMIT = 256 << 20    # bytes per array: larger batches go through the caller's (pageable) buffers

    def staging(self):
        """Five page-locked float32 buffers of the batch shape owned by this engine (x0, obs, true, x, s), or None when the batch is
        larger than STAGING_LIMIT per array or page-locked memory cannot be had.  pds_restore_host moves page-locked buffers chunk by
        chunk under the loop (include/pnp_pds.h); pageable ones cost a staging pass through the driver each way."""
        if getattr(self, "_staging", None) is None:
            self._staging = False
            if 4 * int(np.prod(self.shape)) <= self.STAGING_LIMIT:
                torch = _torch()
                try:
                    self._staging = {k: torch.empty(self.shape, dtype=torch.float32, pin_memory=True) for k in ("x0", "obs", "true", "x", "s")}
                except RuntimeError:
                    self._staging = False
        return self._staging or None

    def stage(self, name: str, src):
        """Copy (and convert to float32, multi-threaded) an array, a tensor or a sequence of per-item arrays into the staging buffer
        `name`; returns its numpy view."""
        torch = _torch()
        dst = self.staging()[name]

        def tensor(a):
            if isinstance(a, torch.Tensor):
                return a
            a = np.asarray(a)
            return torch.from_numpy(a if a.flags.writeable else a.copy())      # from_numpy warns about read-only arrays (npz members)
        if isinstance(src, (list, tuple)):
            if len(src) != self.B:
                raise ValueError("need one array per item")
            put = lambda k: dst[k].copy_(tensor(src[k]).reshape(dst[k].shape))
            if self.B >= 32:
                # a few host threads: under torchrun every rank runs with OMP_NUM_THREADS=1, and the float64 -> float32 conversion
                # of a batch is then the longest host-side step of a short job (copy_ releases the GIL)
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(max_workers=4) as pool:
                    list(pool.map(put, range(self.B)))
            else:
                for k in range(self.B):
                    put(k)
        else:
            dst.copy_(tensor(src).reshape(self.shape))
        return dst.numpy()

    def restore_host(self, x0: np.ndarray, obs: np.ndarray, x_true: np.ndarray | None, n_iter: int, want_s: bool = True,
                     out: np.ndarray | None = None, s_out: np.ndarray | None = None):
        """Whole job through host buffers (H2D + loop + D2H inside the C call).  `out` / `s_out`: optional caller-owned float32
        buffers of the batch shape for x and s (e.g. pinned memory, which the device copies into at full PCIe speed)."""
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32).reshape(self.shape)
        x0, obs, x_true = f(x0), f(obs), f(x_true)
        if out is not None:
            if out.dtype != np.float32 or not out.flags.c_contiguous or out.size != int(np.prod(self.shape)):
                raise ValueError("out must be a C-contiguous float32 array of the batch shape")
            x = out.reshape(self.shape)
        else:
            x = np.empty(self.shape, dtype=np.float32)
        if want_s and s_out is not None:
            if s_out.dtype != np.float32 or not s_out.flags.c_contiguous or s_out.size != int(np.prod(self.shape)):
                raise ValueError("s_out must be a C-contiguous float32 array of the batch shape")
            s = s_out.reshape(self.shape)
        else:
            s = np.empty(self.shape, dtype=np.float32) if want_s else None
        tr = np.zeros((int(n_iter), self.B, TRACE_WIDTH), dtype=np.float64)
        hp = lambda a: C.c_void_p(0) if a is None else a.ctypes.data_as(C.c_void_p)
        _lib.check(self.lib.pds_restore_host(self._h, hp(x0), hp(obs), hp(x_true), int(n_iter), hp(x), hp(s), hp(tr), tr.size,
                                             self._stream()))
        return x, s, tr

    def profile(self, on: bool = True):
        _lib.check(self.lib.pds_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> dict:
        """{category: (milliseconds, launches)} accumulated by CUDA events around each kernel launch."""
        n = len(_lib.PROF_CATS)
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        _lib.check(self.lib.pds_profile_read(self._h, ms, cnt, 1 if reset else 0, self._stream()))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.PROF_CATS)}

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.pds_kernel_launches(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.pds_workspace_bytes(self._h))


def metrics_from_traces(tr: np.ndarray, n: int):
    """c[i] (iteration.py:187) and PSNR (utils_eval.py:4-7) per iteration and item from the raw sums."""
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.sqrt(tr[..., 1]) / np.sqrt(tr[..., 2])
        psnr = 10.0 * np.log10(1.0 / (tr[..., 3] / n))
    return c, psnr


def ssim_from_traces(tr: np.ndarray, C: int, H: int, W: int):
    """Mean SSIM (utils_eval.py:9-12) per iteration and item; NaN where it was not evaluated."""
    positions = (C * (H - 6) * (W - 6)) if C != 1 else (H * (W - 6))
    raw = tr[..., 4]
    if positions <= 0:
        return np.full(raw.shape, np.nan)
    return np.where(raw != 0.0, raw / positions, np.nan)
