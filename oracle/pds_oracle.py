"""CPU oracle for the PnP-PDS hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm
(yodai49/PnP-PDS).  It exists to *check* the CUDA path; nothing under
``pnp-pds_b200/`` may import it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it.

Parity status: PINNED.  Every function below is checked in
``tests/test_oracle_golden.py`` against input/output vectors recorded from the
reference itself (``tests/golden/make_golden.py`` imports ``/root/reference``
with the four shims of SURVEY.md §8c and stores ``tests/golden/*.npz``).
The single exception is ``eval_ssim`` (the reference calls scikit-image, which
is neither in the reference tree nor installed here): SSIM parity is UNPINNED.

Each function cites the reference file:line it follows.  The restatement is
deliberately written in a different form from the reference (spatial periodic
stencils instead of FFTs, a set instead of an O(n^2) list scan, closed-form
loops) so that agreement with the golden vectors is a real check.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# Degradation operators  (reference operators.py:7-79)
# --------------------------------------------------------------------------


def blur_taps(h: np.ndarray, adjoint: bool):
    """Non-zero taps of the blur as (dy, dx, weight) for out[i,j] += w*in[i+dy, j+dx].

    operators.py:7-22 (Phi): wrap-pad by (l//2+1, l//2), FFT-multiply, crop [l:, l:]
      == periodic convolution with the kernel centred at index l//2:
         out[i,j] = sum_{a,b} h[a,b] x[(i + c - a) mod H, (j + c - b) mod W],  c = l//2
    operators.py:24-38 (Phi^T): wrap-pad (c, c), multiply by conj(A), crop [:-l+1]
      == periodic correlation: out[i,j] = sum h[a,b] x[(i + a - c) mod H, (j + b - c) mod W]
    """
    l = h.shape[0]
    c = l // 2
    taps = []
    for a in range(l):
        for b in range(h.shape[1]):
            w = float(h[a, b])
            if w == 0.0:
                continue
            if adjoint:
                taps.append((a - c, b - c, w))
            else:
                taps.append((c - a, c - b, w))
    return taps


def _stencil_periodic(x: np.ndarray, taps) -> np.ndarray:
    out = np.zeros(x.shape, dtype=np.float64)
    xd = x.astype(np.float64, copy=False)
    for dy, dx, w in taps:
        out += w * np.roll(xd, shift=(-dy, -dx), axis=(-2, -1))
    return out


def blur_phi(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """Phi for deg_op='blur' (operators.py:7-22); returns float64 like the reference."""
    return _stencil_periodic(x, blur_taps(h, adjoint=False))


def blur_phi_adj(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """Phi^T for deg_op='blur' (operators.py:24-38)."""
    return _stencil_periodic(x, blur_taps(h, adjoint=True))


def _blur_fft(x: np.ndarray, h: np.ndarray, adjoint: bool) -> np.ndarray:
    """Same periodic operator evaluated with an image-sized real FFT (used by the timed CPU baseline:
    the reference also works in the Fourier domain, operators.py:13-14).  Equal to the stencil form to
    round-off (checked in tests/test_oracle_golden.py)."""
    H, W = x.shape[-2:]
    l = h.shape[0]
    c = l // 2
    k = np.zeros((H, W))
    for a in range(l):
        for b in range(l):
            if h[a, b] != 0.0:
                k[(a - c) % H, (b - c) % W] += h[a, b]      # periodic convolution kernel centred at c
    K = np.fft.rfft2(k)
    if adjoint:
        K = np.conj(K)
    return np.fft.irfft2(np.fft.rfft2(x.astype(np.float64), axes=(-2, -1)) * K, s=(H, W), axes=(-2, -1))


def sampling_mask(H: int, W: int, r: float) -> np.ndarray:
    """uint8 (H, W) keep-mask of the random_sampling operator (operators.py:40-58).

    The reference zeroes the first round(H*W*(1-r)) entries of
    RandomState(1234).permutation(H*W) in row-major flat order, identically for
    every channel, image and call.  Integer work: must be bit-exact.
    """
    n_drop = round(H * W * (1 - r))
    q = np.random.RandomState(seed=1234).permutation(H * W)[:n_drop]
    m = np.ones(H * W, dtype=np.uint8)
    m[q] = 0
    return m.reshape(H, W)


def sample(x: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Phi = Phi^T for random_sampling (operators.py:40-58, 65, 73)."""
    return x * mask.astype(x.dtype)


def make_operators(deg_op: str, h: np.ndarray | None, r: float, shape_hw=None, fft: bool = False):
    """(phi, adj_phi) closures like operators.get_observation_operators (operators.py:60-79)."""
    if deg_op == "blur":
        if fft:
            return (lambda x: _blur_fft(x, h, False)), (lambda x: _blur_fft(x, h, True))
        return (lambda x: blur_phi(x, h)), (lambda x: blur_phi_adj(x, h))
    if deg_op == "random_sampling":
        cache = {}

        def f(x):
            key = x.shape[-2:]
            if key not in cache:
                cache[key] = sampling_mask(key[0], key[1], r)
            return sample(x, cache[key])

        return f, f
    if deg_op == "Id":
        return (lambda x: x), (lambda x: x)
    raise ValueError(f"unknown deg_op {deg_op!r}")


# --------------------------------------------------------------------------
# Proximal maps  (reference operators.py:94-115)
# --------------------------------------------------------------------------


def l2_ball_radius(n: int, alpha_n: float, gaussian_nl: float, sp_nl: float, r: float = 1.0) -> float:
    """epsilon of operators.py:104."""
    return float(np.sqrt(n * (1 - sp_nl)) * r * alpha_n * gaussian_nl)


def l1_ball_radius(n: int, alpha_s: float, sp_nl: float, r: float = 1.0) -> float:
    """eta of operators.py:96."""
    return float(alpha_s * n * sp_nl * r * 0.5)


def proj_l2_ball(x, alpha_n, gaussian_nl, sp_nl, x_0, r=1):
    """operators.py:102-108: projection onto {z : ||z - x_0||_2 <= eps}."""
    eps = l2_ball_radius(x.size, alpha_n, gaussian_nl, sp_nl, r)
    d = x - x_0
    nrm = np.sqrt(np.sum(np.square(d, dtype=np.float64)))
    if nrm > eps:
        return x_0 + d * (eps / nrm)
    return np.array(x, copy=True)


def l1_threshold(absz: np.ndarray, eta: float) -> float:
    """tau of operators.py:98: max(0, max_k (cumsum(sort_desc|z|)_k - eta)/k)."""
    a = np.sort(absz.reshape(-1).astype(np.float64))[::-1]
    cs = np.cumsum(a)
    k = np.arange(1, a.size + 1, dtype=np.float64)
    return float(max(np.max((cs - eta) / k), 0.0))


def proj_l1_ball(x, alpha_s, sp_nl, r=1):
    """operators.py:94-100: soft-threshold by the l1-ball threshold."""
    eta = l1_ball_radius(x.size, alpha_s, sp_nl, r)
    tau = l1_threshold(np.abs(x), eta)
    return np.sign(x) * np.maximum(np.abs(x) - tau, 0.0)


def prox_gkl(x, gamma, alpha, x_0):
    """operators.py:114-115: positive root of p^2 - (x - gamma*alpha) p - gamma*x_0 = 0."""
    q = x - gamma * alpha
    return 0.5 * (q + np.sqrt(q * q + 4.0 * gamma * x_0))


def grad_x_l2(x, s, phi, adj_phi, x_0):
    """operators.py:88-89."""
    return 2 * adj_phi(phi(x) + s - x_0)


def grad_s_l2(x, s, phi, x_0):
    """operators.py:91-92."""
    return phi(x) + s - x_0


# --------------------------------------------------------------------------
# Total-variation pieces of the TV baselines  (reference operators.py:110-137; colour images only — the
# reference hard-codes three channels)
# --------------------------------------------------------------------------


def tv_D(x: np.ndarray) -> np.ndarray:
    """operators.py:117-125: forward differences, (3,H,W) -> (6,H,W) = [vertical (axis 1); horizontal (axis 2)], last row /
    column zero."""
    C, H, W = x.shape
    out = np.zeros((2 * C, H, W))
    out[:C, :H - 1, :] = x[:, 1:, :] - x[:, :-1, :]
    out[C:, :, :W - 1] = x[:, :, 1:] - x[:, :, :-1]
    return out


def tv_DT(y: np.ndarray) -> np.ndarray:
    """operators.py:127-137.  The exact adjoint of tv_D except in the last row / column, where the reference returns
    +y[last] instead of +y[last-1].  Restated as written."""
    C = y.shape[0] // 2
    v, h = y[:C], y[C:]
    H, W = v.shape[1], v.shape[2]
    xv = np.empty(v.shape)
    xv[:, 0, :] = -v[:, 0, :]
    xv[:, 1:H - 1, :] = -v[:, 1:H - 1, :] + v[:, 0:H - 2, :]
    xv[:, H - 1, :] = v[:, H - 1, :]
    xh = np.empty(h.shape)
    xh[:, :, 0] = -h[:, :, 0]
    xh[:, :, 1:W - 1] = -h[:, :, 1:W - 1] + h[:, :, 0:W - 2]
    xh[:, :, W - 1] = h[:, :, W - 1]
    return xv + xh


def prox_l12(x: np.ndarray, gamma: float) -> np.ndarray:
    """operators.py:110-112: group soft-threshold over axis 0 (all six difference channels of a pixel together)."""
    nrm = np.sqrt(np.sum(x * x, 0))
    with np.errstate(divide="ignore", invalid="ignore"):
        val = gamma / nrm
    return np.fmax(1 - val, 0) * x


# --------------------------------------------------------------------------
# DnCNN denoiser  (reference models/denoiser.py:34-46, models/basic_models.py:25-38,
#                  models/network_dncnn.py:42-77)
# --------------------------------------------------------------------------


def conv3x3(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """3x3, stride 1, zero padding 1 cross-correlation (torch.nn.Conv2d semantics), float32.

    x: (Cin, H, W); w: (Cout, Cin, 3, 3); b: (Cout,)
    """
    cin, H, W = x.shape
    cout = w.shape[0]
    xp = np.zeros((cin, H + 2, W + 2), dtype=np.float32)
    xp[:, 1:-1, 1:-1] = x
    cols = np.empty((cin * 9, H * W), dtype=np.float32)
    i = 0
    for c in range(cin):
        for dy in range(3):
            for dx in range(3):
                cols[i] = xp[c, dy:dy + H, dx:dx + W].reshape(-1)
                i += 1
    out = w.reshape(cout, cin * 9).astype(np.float32) @ cols
    out += b.astype(np.float32)[:, None]
    return out.reshape(cout, H, W)


def dncnn_forward(layers, x: np.ndarray, slope: float = 0.01, residual_sign: float = 1.0,
                  clamp: bool = True) -> np.ndarray:
    """Denoiser forward on one image (C,H,W) or (H,W) -> float32 same shape.

    simple_CNN (basic_models.py:25-38) wrapped by apply_model (denoiser.py:34-46):
      clamp(x,0,1) -> conv+LeakyReLU(0.01) x19 -> conv -> + clamped input -> clamp(0,1)
    KAIR DnCNN (network_dncnn.py:75-77): slope=0, residual_sign=-1 (x - model(x)), clamp=False.
    layers: list of (weight (Cout,Cin,3,3), bias (Cout,)) float32.
    """
    squeeze = x.ndim == 2
    a = np.asarray(x, dtype=np.float32)
    if squeeze:
        a = a[None]
    if clamp:
        a = np.clip(a, 0.0, 1.0)
    x_in = a
    for i, (w, b) in enumerate(layers):
        a = conv3x3(a, w, b)
        if i != len(layers) - 1:
            a = np.where(a >= 0, a, np.float32(slope) * a).astype(np.float32)
    if residual_sign > 0:
        out = a + x_in
    else:
        out = x_in - a
    if clamp:
        out = np.clip(out, 0.0, 1.0)
    out = out.astype(np.float32)
    return out[0] if squeeze else out


def dncnn_forward_torch(layers, x: np.ndarray, slope: float = 0.01, residual_sign: float = 1.0,
                        clamp: bool = True) -> np.ndarray:
    """Same network through torch's CPU conv2d with all host threads — what the reference itself runs
    on a CPU-only box (denoiser.py:34-46).  Used for the timed CPU baseline; agrees with
    dncnn_forward to fp32 summation order."""
    import torch
    import torch.nn.functional as F
    squeeze = x.ndim == 2
    a = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    a = a[None, None] if squeeze else a[None]
    with torch.no_grad():
        if clamp:
            a = a.clamp(0, 1)
        x_in = a
        for i, (w, b) in enumerate(layers):
            a = F.conv2d(a, torch.from_numpy(w), torch.from_numpy(b), padding=1)
            if i != len(layers) - 1:
                a = F.leaky_relu(a, slope)
        out = a + x_in if residual_sign > 0 else x_in - a
        if clamp:
            out = out.clamp(0, 1)
    out = out[0].numpy()
    return out[0] if squeeze else out


# --------------------------------------------------------------------------
# Metrics  (reference utils/utils_eval.py, iteration.py:187)
# --------------------------------------------------------------------------


def eval_psnr(im1, im2):
    """utils_eval.py:4-7."""
    mse = np.mean((im1.astype(np.float64) - im2.astype(np.float64)) ** 2)
    return 10 * np.log10(1.0 / mse)


def _uniform_filter_valid(a: np.ndarray, win: int, axes) -> np.ndarray:
    """Mean over a win-wide window along each axis in `axes`, 'valid' region only."""
    out = a.astype(np.float64)
    for ax in axes:
        cs = np.cumsum(np.insert(out, 0, 0.0, axis=ax), axis=ax)
        n = out.shape[ax]
        hi = np.take(cs, np.arange(win, n + 1), axis=ax)
        lo = np.take(cs, np.arange(0, n - win + 1), axis=ax)
        out = (hi - lo) / win
    return out


def _ssim_single(a: np.ndarray, b: np.ndarray, data_range: float, win: int = 7) -> float:
    """skimage.metrics.structural_similarity for one n-D channel, default settings
    (uniform 7-tap window, K1=0.01, K2=0.03, sample covariance, mean over the
    region cropped by (win-1)//2 — identical to the 'valid' window positions)."""
    axes = tuple(range(a.ndim))
    NP = win ** a.ndim
    cov_norm = NP / (NP - 1.0)
    ux = _uniform_filter_valid(a, win, axes)
    uy = _uniform_filter_valid(b, win, axes)
    uxx = _uniform_filter_valid(a * a, win, axes)
    uyy = _uniform_filter_valid(b * b, win, axes)
    uxy = _uniform_filter_valid(a * b, win, axes)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (0.01 * data_range) ** 2
    C2 = (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return float(np.mean(S))


def eval_ssim(im1, im2):
    """utils_eval.py:9-12 — PARITY UNPINNED (scikit-image absent; restated from its
    documented defaults).  channel_axis=0: a gray (H,W) image is treated as H
    one-dimensional channels; colour (3,H,W) as three 2-D channels."""
    data_range = float(im2.max() - im2.min())
    a = im1.astype(np.float64)
    b = im2.astype(np.float64)
    vals = [_ssim_single(a[c], b[c], data_range) for c in range(a.shape[0])]
    return float(np.mean(vals))


# --------------------------------------------------------------------------
# Observation synthesis  (reference main.py:49-64, utils/utils_noise.py)
# --------------------------------------------------------------------------


def add_gaussian_noise(img, noise_level, op):
    """utils_noise.py:33-36 (seed 1234, noise passed through `op`)."""
    np.random.seed(1234)
    return img + op(noise_level * np.random.randn(*img.shape))


def apply_poisson_noise(img, alpha):
    """utils_noise.py:38-41."""
    np.random.seed(1234)
    return np.random.poisson(img * alpha)


def add_salt_and_pepper_noise(img, noise_level, op):
    """utils_noise.py:3-31.  Both coordinates are drawn from randint(0, H) (H=shape[-2]);
    a draw is kept only when the pixel is observed (op(ones)==1) and not yet used;
    the reference's ``i = i - 1`` retry has no effect, so 2*noise_cnt draws are made in
    total.  The first noise_cnt kept points become 0, the remainder 1."""
    H = img.shape[-2]
    noise_cnt = int(img.shape[-2] * img.shape[-1] * noise_level / 2)
    target = op(np.ones([img.shape[-2], img.shape[-1]]))
    np.random.seed(1234)
    xs, ys, seen = [], [], set()
    for _ in range(noise_cnt * 2):
        x = np.random.randint(0, H)
        y = np.random.randint(0, H)
        key = x * H + y
        if target[x][y] == 1 and key not in seen:
            seen.add(key)
            xs.append(x)
            ys.append(y)
    xs = np.asarray(xs, dtype=np.int64)
    ys = np.asarray(ys, dtype=np.int64)
    out = np.copy(img)
    planes = [out] if out.ndim == 2 else [out[i] for i in range(3)]
    for p in planes:
        p[(xs[:noise_cnt], ys[:noise_cnt])] = 0
        p[(xs[noise_cnt:], ys[noise_cnt:])] = 1
    return out


def synthesize_observation(img_true, deg_op, h, r, gaussian_nl, sp_nl, poisson_noise, poisson_alpha):
    """main.py:49-64: returns (x_0, img_obsrv)."""
    phi, _ = make_operators(deg_op, h, r)
    ident = (lambda z: z)
    noise_op = phi if deg_op == "random_sampling" else ident
    obs = phi(img_true)
    obs = add_gaussian_noise(obs, gaussian_nl, noise_op)
    if poisson_noise:
        obs = apply_poisson_noise(obs, poisson_alpha)
    obs = add_salt_and_pepper_noise(obs, sp_nl, noise_op)
    x0 = np.copy(obs)
    if poisson_noise:
        x0 = x0 / poisson_alpha
    return x0, obs


def synthetic_image(b: int, C: int, H: int, W: int) -> np.ndarray:
    """Frozen synthetic test image: smooth pattern + seeded texture,
    clipped to [0.05, 0.95] so Poisson rates stay positive.  float32, (H,W) or (C,H,W)."""
    v, u = np.meshgrid(np.arange(W) / W, np.arange(H) / H)
    base = 0.5 + 0.3 * np.sin(2 * np.pi * 2 * u) * np.cos(2 * np.pi * 3 * v)
    rng = np.random.default_rng(seed=b)
    planes = []
    for c in range(C):
        U = rng.random((H, W))
        planes.append(np.clip(base + 0.05 * c + 0.1 * (U - 0.5), 0.05, 0.95))
    img = np.stack(planes).astype(np.float32)
    return img[0] if C == 1 else img


# --------------------------------------------------------------------------
# The iteration  (reference iteration.py:10-196)
# --------------------------------------------------------------------------

METHOD_ALIASES = {
    "ours-A": "A-Proposed", "ours-B": "B-Proposed", "ours-C": "C-Proposed",
    "comparisonA-1": "A-PnPFBS-DnCNN", "comparisonA-6": "A-RED-DnCNN",
    "comparisonC-2": "C-PnPADMM-DnCNN", "comparisonC-3": "C-RED-DnCNN", "comparisonA-4": "A-PDS-TV",
}


def pds_iterations(x_0, x_obsrv, x_true, phi, adj_phi, denoise, gamma1, gamma2, alpha_s, alpha_n,
                   myLambda, gaussian_nl, sp_nl, poisson_alpha, max_iter, method="A-Proposed", r=1,
                   m1=15, m2=15, gammaInADMMStep1=0.1, snapshots=(), y0=None, s0=None):
    """Restatement of iteration.test_iter for the DnCNN-based methods.

    `denoise` is a callable image->float32 image (the Denoiser.denoise of denoiser.py:14-16).
    Returns (x, s+0.5, c, psnr, snaps) with snaps[i] = dict(x=, y=, s=) for i in `snapshots`
    (1-based iteration numbers).  SSIM is left to the caller.
    """
    method = METHOD_ALIASES.get(method, method)
    x = x_0
    y = np.zeros(x_0.shape) if y0 is None else y0      # y0/s0: resume from a previous call (bench CPU baseline)
    s = np.zeros(x_0.shape) if s0 is None else s0
    z = np.zeros(x_0.shape)
    d = np.zeros(x_0.shape)
    y1 = np.concatenate([np.zeros(x_0.shape), np.zeros(x_0.shape)], 0)     # iteration.py:25 (TV baselines)
    c = np.zeros(max_iter)
    psnr = np.zeros(max_iter)
    snaps = {}
    n = x_0.size
    for i in range(max_iter):
        x_prev, s_prev = x, s
        if method == "A-Proposed":                                   # iteration.py:48-52
            x = denoise(x - gamma1 * adj_phi(y))
            w = y + gamma2 * phi(2 * x - x_prev)
            eps = l2_ball_radius(n, alpha_n, gaussian_nl, sp_nl, 1)      # r NOT passed (quirk Q1)
            t = w - gamma2 * x_obsrv
            nt = np.sqrt(np.sum(t * t))
            y = t * max(0.0, 1.0 - gamma2 * eps / nt) if nt > 0 else t * 0.0
        elif method == "B-Proposed":                                 # iteration.py:53-58
            x = denoise(x - gamma1 * adj_phi(y))
            s = proj_l1_ball(s - gamma1 * y, alpha_s, sp_nl, r)
            w = y + gamma2 * (phi(2 * x - x_prev) + 2 * s - s_prev)
            eps = l2_ball_radius(n, alpha_n, gaussian_nl, sp_nl, r)
            t = w - gamma2 * x_obsrv
            nt = np.sqrt(np.sum(t * t))
            y = t * max(0.0, 1.0 - gamma2 * eps / nt) if nt > 0 else t * 0.0
        elif method == "C-Proposed":                                 # iteration.py:59-63
            x = denoise(x - gamma1 * adj_phi(y))
            w = y + gamma2 * phi(2 * x - x_prev)
            la = myLambda * poisson_alpha
            y = 0.5 * (w + la - np.sqrt((w - la) ** 2 + 4.0 * myLambda * gamma2 * x_obsrv))
        elif method == "A-PnPFBS-DnCNN":                             # iteration.py:71-73
            x = denoise(x - gamma1 * myLambda * 0.5 * grad_x_l2(x, np.zeros(x.shape), phi, adj_phi, x_obsrv))
        elif method == "A-RED-DnCNN":                                # iteration.py:100-105
            dx = denoise(x)
            mu = 2 / (1 / gamma1 ** 2 + myLambda)
            x = x_prev - mu * ((1 / gamma1 ** 2) * adj_phi(phi(x_prev) - x_obsrv) + myLambda * (x_prev - dx))
        elif method == "comparisonB-2":                              # iteration.py:127-132, admm.py:30-44
            xx = np.ones(s.shape)
            for _ in range(m1):
                xx = denoise(xx - (1 / gamma1) * adj_phi(phi(xx) + s - z + y))
            x = xx
            ss = np.ones(x.shape)
            for _ in range(m2):
                ss = proj_l1_ball(ss - (1 / gamma1) * (phi(x) + ss - z + y), alpha_s, sp_nl)
            s = ss
            z = proj_l2_ball(phi(x) + s + y, alpha_n, gaussian_nl, sp_nl, x_obsrv)
            y = y + phi(x) + s - z
        elif method == "comparisonB-4":                              # iteration.py:141-145
            x = x_prev - gamma1 * (myLambda * adj_phi(phi(x_prev) + s - x_obsrv) + (x_prev - denoise(x)))
            s = proj_l1_ball(s - gamma1 * grad_s_l2(x, s, phi, x_obsrv), alpha_s, sp_nl)
        elif method == "comparisonB-5":                              # iteration.py:146-149
            x = denoise(x - gamma1 * grad_x_l2(x, s, phi, adj_phi, x_obsrv))
            s = proj_l1_ball(s - gamma1 * grad_s_l2(x, s, phi, x_obsrv), alpha_s, sp_nl)
        elif method == "A-PDS-TV":                                   # iteration.py:88-94 (y plays the role of y2_n)
            x = x - gamma1 * (tv_DT(y1) + adj_phi(y))
            y1 = y1 + gamma2 * tv_D(2 * x - x_prev)
            y1 = y1 - gamma2 * prox_l12(y1 / gamma2, 1 / gamma2)
            y = y + gamma2 * phi(2 * x - x_prev)
            y = y - gamma2 * proj_l2_ball(y / gamma2, alpha_n, gaussian_nl, sp_nl, x_obsrv)
        elif method == "A-FBS-TV":                                   # iteration.py:95-99
            x = x - gamma1 * (adj_phi(phi(x) - x_obsrv) + tv_DT(y1))
            y1 = y1 + gamma2 * tv_D(2 * x - x_prev)
            y1 = y1 - gamma2 * prox_l12(y1 / gamma2, 1 / gamma2)
        elif method == "comparisonB-3":                              # iteration.py:133-140 (both projections WITHOUT r)
            x = x - gamma1 * (tv_DT(y1) + adj_phi(y))
            s = proj_l1_ball(s - gamma1 * y, alpha_s, sp_nl)
            y1 = y1 + gamma2 * tv_D(2 * x - x_prev)
            y1 = y1 - gamma2 * prox_l12(y1 / gamma2, 1 / gamma2)
            y = y + gamma2 * (phi(2 * x - x_prev) + 2 * s - s_prev)
            y = y - gamma2 * proj_l2_ball(y / gamma2, alpha_n, gaussian_nl, sp_nl, x_obsrv)
        elif method in ("C-PnPADMM-DnCNN", "C-RED-DnCNN"):          # iteration.py:161-172, admm.py:4-28
            xx = np.ones(d.shape)
            ones_adj = adj_phi(np.ones(xx.shape))
            for _ in range(m1):
                grad = (-adj_phi(x_obsrv / (poisson_alpha * phi(xx))) / poisson_alpha
                        + ones_adj / poisson_alpha + myLambda * (xx - z + d))
                xx = xx - gammaInADMMStep1 * grad
            x = xx
            if method == "C-PnPADMM-DnCNN":
                z = denoise(x + d)
            else:
                z_str = x + d
                zz = z
                for _ in range(m2):
                    zz = denoise(zz)
                    zz = 1 / (myLambda + gamma1) * (gamma1 * zz + myLambda * z_str)
                z = zz
            d = d + x - z
        else:
            raise ValueError(f"oracle: unsupported method {method!r}")
        c[i] = np.linalg.norm((x - x_prev).ravel()) / np.linalg.norm(np.asarray(x_prev).ravel())
        psnr[i] = eval_psnr(x_true, x)
        if (i + 1) in snapshots:
            snaps[i + 1] = dict(x=np.array(x, copy=True), y=np.array(y, copy=True), s=np.array(s, copy=True))
    return x, s + 0.5, c, psnr, snaps
