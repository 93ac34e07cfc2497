"""Experiment driver with the reference's interface (main.py:16-123) feeding the batched GPU engine.

  test_all_images(experimental_settings_arg, method_arg, configs_arg) -> datas   (main.py:16)
  eval_restoration(**kwargs) -> psnr                                               (ideas/param_memo.py:3)
  grid_search(images, grid, ...)   sharded over the GPUs of one box                (main.py:125-159 sweep)

Differences from the reference, by design: `config/setup.json` is read when a function is called, not
at import; all images of one shape are restored as ONE resident batch instead of sequentially; paths
are joined portably (the reference hard-codes Windows back-slashes, main.py:76,91,119).
"""
from __future__ import annotations

import datetime
import glob
import json
import os

import numpy as np

from . import iteration
from .operators import get_observation_operators
from .parallel import dist_info, gather_rows, shard_range
from .utils.utils_eval import eval_psnr, eval_ssim
from .utils.utils_method_master import get_algorithm_denoiser
from .utils.utils_noise import add_gaussian_noise, add_salt_and_pepper_noise, apply_poisson_noise
from .utils.utils_parse_args import parse_args_configs, parse_args_exp, parse_args_method
from .utils.utils_unparse_args import unparse_args_configs, unparse_args_exp, unparse_args_method


def load_config(path="config/setup.json"):
    """README.md:6-14 schema: path_test, path_result, pattern_red, root_folder."""
    with open(path, "r") as f:
        return json.load(f)


def synthesize_observation(img_true, phi, deg_op, gaussian_nl, sp_nl, poisson_noise, poisson_alpha):
    """main.py:49-64: Phi(x) -> + Gaussian -> Poisson -> salt&pepper; x_0 = copy (/alpha if Poisson).
    The noise realisation is the same for every image (constant seeds)."""
    ident = (lambda z: z)
    noise_op = phi if deg_op == "random_sampling" else ident
    obs = phi(img_true)
    obs = add_gaussian_noise(obs, gaussian_nl, noise_op)
    if poisson_noise:
        obs = apply_poisson_noise(obs, poisson_alpha)
    obs = add_salt_and_pepper_noise(obs, sp_nl, noise_op)
    x_0 = np.copy(obs)
    if poisson_noise:
        x_0 = x_0 / poisson_alpha
    return x_0, obs


def read_image(path, ch):
    """main.py:41-48"""
    import cv2
    img = np.asarray(cv2.imread(path), dtype="float32") / 255.0
    if ch == 1:
        return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    return np.moveaxis(img, -1, 0)


def restore_images(images, names, experimental_settings_arg=None, method_arg=None, configs_arg=None, path_kernel=None,
                   path_prox=None, conv_engine="tcgen05"):
    """Core of test_all_images on in-memory images (list of (H,W) or (3,H,W) float arrays)."""
    gaussian_nl, sp_nl, poisson_noise, poisson_alpha, deg_op, r = parse_args_exp(experimental_settings_arg or {})
    (method, architecture, max_iter, gamma1, gamma2, alpha_n, alpha_s, myLambda, m1, m2,
     gammaInADMMStep1) = parse_args_method(method_arg or {})
    ch, add_timestamp, result_output = parse_args_configs(configs_arg or {})
    phi, adj_phi = get_observation_operators(deg_op, path_kernel, r)
    n_img = len(images)
    psnr, ssim, cpu_time = np.zeros(n_img), np.zeros(n_img), np.zeros(n_img)
    results = {}
    prm = dict(gamma1=gamma1, gamma2=gamma2, alpha_s=alpha_s, alpha_n=alpha_n, myLambda=myLambda, gaussian_nl=gaussian_nl,
               sp_nl=sp_nl, poisson_alpha=poisson_alpha, r=r)
    by_shape = {}
    for i, im in enumerate(images):
        by_shape.setdefault(tuple(np.shape(im)), []).append(i)
    for shape, idxs in by_shape.items():
        trues = np.stack([np.asarray(images[i]) for i in idxs])
        pairs = [synthesize_observation(trues[k], phi, deg_op, gaussian_nl, sp_nl, poisson_noise, poisson_alpha)
                 for k in range(len(idxs))]
        x0 = np.stack([p[0] for p in pairs])
        obs = np.stack([p[1] for p in pairs])
        res = iteration.run_batch(x0, obs, trues, phi, adj_phi, prm, path_prox, max_iter, method, ch, conv_engine=conv_engine)
        for k, i in enumerate(idxs):
            img_obsrv = obs[k] / poisson_alpha if poisson_noise else obs[k]
            psnr[i], ssim[i], cpu_time[i] = res["psnr"][-1, k], res["ssim"][-1, k], res["time_per_iter"]
            results[i] = {
                "filename": names[i], "c_evolution": res["c"][:, k], "PSNR_evolution": res["psnr"][:, k],
                "SSIM_evolution": res["ssim"][:, k], "GROUND_TRUTH": trues[k], "OBSERVATION": img_obsrv, "RESULT": res["x"][k],
                "REMOVED_SPARSE": res["s"][k].astype(np.float64) + 0.5, "PSNR": psnr[i], "SSIM": ssim[i], "CPU_time": cpu_time[i],
                "PSNR_observation": eval_psnr(trues[k], img_obsrv), "SSIM_observation": eval_ssim(trues[k], img_obsrv),
            }
    algorithm, denoiser = get_algorithm_denoiser(method)
    summary = {"Average_PSNR": np.mean(psnr), "PSNR": psnr, "Average_SSIM": np.mean(ssim), "SSIM": ssim,
               "Average_time": np.average(cpu_time), "Cpu_time": cpu_time, "algorithm": algorithm, "denoiser": denoiser}
    return {
        "experimental_settings": unparse_args_exp(gaussian_nl, sp_nl, poisson_noise, poisson_alpha, deg_op, r),
        "method": unparse_args_method(method, architecture, max_iter, gamma1, gamma2, alpha_n, alpha_s, myLambda, m1, m2,
                                      gammaInADMMStep1),
        "configs": unparse_args_configs(ch, add_timestamp, result_output),
        "results": results, "summary": summary,
    }


def test_all_images(experimental_settings_arg={}, method_arg={}, configs_arg={}, config=None):
    """main.py:16-123"""
    config = config or load_config()
    _, architecture, *_ = parse_args_method(method_arg)
    ch, add_timestamp, _ = parse_args_configs(configs_arg)
    _, _, _, _, deg_op, _ = parse_args_exp(experimental_settings_arg)
    method = parse_args_method(method_arg)[0]
    gaussian_nl = parse_args_exp(experimental_settings_arg)[0]
    path_kernel = os.path.join(config["root_folder"], "blur_models", "blur_1.mat")       # always blur_1 (main.py:27)
    path_prox = os.path.join(config["root_folder"], "nn", architecture + ".pth")
    path_images = sorted(glob.glob(os.path.join(config["path_test"], config["pattern_red"])))
    images = [read_image(p, ch) for p in path_images]
    names = [os.path.basename(p) for p in path_images]
    datas = restore_images(images, names, experimental_settings_arg, method_arg, configs_arg, path_kernel, path_prox)
    path_result = config.get("path_result")
    if path_result and os.path.isdir(path_result) and names:
        base = method + "_" + deg_op + "_" + str(gaussian_nl).ljust(5, "0") + "_(" + names[-1] + ")"
        if add_timestamp:
            base += "_" + datetime.datetime.now().strftime("%Y%m%d-%H%M%S-%f")
        np.save(os.path.join(path_result, "DATA_" + base), datas)
    s = datas["summary"]
    print(datetime.datetime.now().strftime("%Y/%m/%d %H:%M:%S") + "  Average_PSNR:" + str(np.round(s["Average_PSNR"], 3))
          + "  Average_SSIM:" + str(np.round(s["Average_SSIM"], 3)) + "    Algorithm:" + method + "   Observation:" + deg_op
          + "   Gaussian noise level:" + str(gaussian_nl).ljust(5, "0"))
    return datas


test_all_images.__test__ = False


def eval_restoration(gaussian_nl=0.01, sp_nl=0.0, poisson_noise=False, poisson_alpha=300, max_iter=10, gamma1=1, gamma2=1,
                     r=0.8, alpha_n=1, alpha_s=1, myLambda=1, result_output=False, architecture="DnCNN_nobn_nch_3_nlev_0.01",
                     deg_op="blur", method="ours-A", ch=3, m1=15, m2=15, gammaInADMMStep1=0.1, config=None):
    """Legacy entry point recorded in ideas/param_memo.py:3-99: keyword arguments -> average PSNR."""
    datas = test_all_images(
        dict(gaussian_nl=gaussian_nl, sp_nl=sp_nl, poisson_noise=poisson_noise, poisson_alpha=poisson_alpha, deg_op=deg_op, r=r),
        dict(method=method, architecture=architecture, max_iter=max_iter, gamma1=gamma1, gamma2=gamma2, alpha_n=alpha_n,
             alpha_s=alpha_s, myLambda=myLambda, m1=m1, m2=m2, gammaInADMMStep1=gammaInADMMStep1),
        dict(ch=ch, add_timestamp=True, result_output=result_output), config=config)
    return datas["summary"]["Average_PSNR"]


def grid_search(images, grid, experimental_settings, method_common, ch, path_kernel, path_prox, batch_size=64,
                conv_engine="tcgen05", timings=None):
    """Hyper-parameter sweep of main.main() (main.py:125-159) over `images` x `grid`, sharded over the ranks
    of a torchrun job.  grid: list of dicts overriding gamma1/gamma2/alpha_n/alpha_s/myLambda.
    Work item = (image i, grid point g); items are split in contiguous blocks over ranks, each rank batches
    its items (mixed grid points in one batch: per-item parameter vectors) and the final
    (psnr, ssim, c_last) rows are all-gathered once.  Returns array [n_images, n_grid, 3] on every rank.
    timings: optional dict that receives this rank's wall seconds per phase (synthesis, restore, gather)."""
    import time
    t_syn = t_run = 0.0
    gaussian_nl, sp_nl, poisson_noise, poisson_alpha, deg_op, r = parse_args_exp(experimental_settings)
    method, _, max_iter, *_ = parse_args_method(method_common)
    rank, local_rank, world = dist_info()
    phi, adj_phi = get_observation_operators(deg_op, path_kernel, r)
    n_img, n_grid = len(images), len(grid)
    n_items = n_img * n_grid
    lo, hi = shard_range(n_items, rank, world)
    obs_cache = {}
    rows = np.zeros((hi - lo, 3))
    for b0 in range(lo, hi, batch_size):
        ids = list(range(b0, min(hi, b0 + batch_size)))
        x0s, obss, trues, prms = [], [], [], []
        for it in ids:
            i, g = divmod(it, n_grid)
            if i not in obs_cache:
                t0 = time.perf_counter()
                obs_cache[i] = synthesize_observation(np.asarray(images[i]), phi, deg_op, gaussian_nl, sp_nl, poisson_noise,
                                                      poisson_alpha)
                t_syn += time.perf_counter() - t0
            x0s.append(obs_cache[i][0]); obss.append(obs_cache[i][1]); trues.append(np.asarray(images[i]))
            base = dict(zip(("method", "architecture", "max_iter", "gamma1", "gamma2", "alpha_n", "alpha_s", "myLambda", "m1", "m2",
                             "gammaInADMMStep1"), parse_args_method({**method_common, **grid[g]})))
            prms.append(dict(gamma1=base["gamma1"], gamma2=base["gamma2"], alpha_s=base["alpha_s"], alpha_n=base["alpha_n"],
                             myLambda=base["myLambda"], gaussian_nl=gaussian_nl, sp_nl=sp_nl, poisson_alpha=poisson_alpha, r=r))
        t0 = time.perf_counter()
        res = iteration.run_batch(x0s, obss, trues, phi, adj_phi, prms, path_prox, max_iter, method,   # per-item arrays: no stacking pass
                                  ch, conv_engine=conv_engine, device=local_rank if world > 1 else None)
        t_run += time.perf_counter() - t0
        for k, it in enumerate(ids):
            rows[it - lo] = (res["psnr"][-1, k], res["ssim"][-1, k], res["c"][-1, k])
    t0 = time.perf_counter()
    table = gather_rows(rows, n_items).reshape(n_img, n_grid, 3)
    if timings is not None:
        timings.update(synthesis_s=t_syn, restore_s=t_run, collective_s=time.perf_counter() - t0, items_on_rank=hi - lo)
    return table


def sweep_experiments():
    """The experiment list of the reference's sweep driver (main.py:130-153): 5 noise levels x {blur, random_sampling} x
    {A-Proposed, A-PDS-TV: alpha_n = 0.82 ... 1.00; A-PnPFBS-DnCNN, A-RED-DnCNN: myLambda = 0.2 ... 1.99}."""
    out = []
    for nl in (0.0025, 0.005, 0.01, 0.02, 0.04):
        for obs in ("blur", "random_sampling"):
            max_iter = 1200 if obs == "blur" else 3000
            settings = {"gaussian_nl": nl, "sp_nl": 0, "poisson_noise": False, "deg_op": obs, "r": 0.8}
            for method in ("A-Proposed", "A-PDS-TV"):
                for i in range(10):
                    out.append({"settings": settings, "configs": {},
                                "method": {"method": method, "max_iter": max_iter, "gamma1": 0.125 if method == "A-PDS-TV" else 0.99,
                                           "gamma2": 0.99, "alpha_n": 0.8 + (i + 1) * 0.02}})
            for method in ("A-PnPFBS-DnCNN", "A-RED-DnCNN"):
                for i in range(10):
                    lam = (i + 1) * 0.2
                    out.append({"settings": settings, "configs": {},
                                "method": {"method": method, "max_iter": max_iter, "gamma1": 1, "myLambda": 1.99 if lam == 2 else lam}})
    return out


def main(config=None, experiments=None):
    """main.py:125-159: run the sweep, one SUMMARY(<timestamp>).txt line per experiment (utils_textfile)."""
    from .utils.utils_textfile import add_footer_textfile, touch_textfile, write_textfile
    config = config or load_config()
    filepath = os.path.join(config["path_result"], "SUMMARY(" + datetime.datetime.now().strftime("%Y%m%d %H%M%S %f") + ").txt")
    touch_textfile(filepath)
    data = None
    for e in (sweep_experiments() if experiments is None else experiments):
        data = test_all_images(e["settings"], e["method"], e["configs"], config=config)
        write_textfile(filepath, data)
    if data is not None:
        add_footer_textfile(filepath, data)
    return filepath


if __name__ == "__main__":
    main()
