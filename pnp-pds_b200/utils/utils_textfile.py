"""SUMMARY(...).txt writer of the sweep driver — same file format as the reference (utils/utils_textfile.py:1-74):
one header line, one comma-separated line per experiment (settings, averages, then per-image PSNR / SSIM / observed PSNR /
observed SSIM, each value followed by a comma), and a footer line with the file names."""
from __future__ import annotations

# (column title, section of the `datas` dict, key)
_COLUMNS = (
    ("Observation", "experimental_settings", "deg_op"),
    ("Gaussian_noise", "experimental_settings", "gaussian_nl"),
    ("Poisson_alpha", "experimental_settings", "poisson_alpha"),
    ("method", "method", "method"),
    ("algorithm", "summary", "algorithm"),
    ("denoiser", "summary", "denoiser"),
    ("PSNR", "summary", "Average_PSNR"),
    ("SSIM", "summary", "Average_SSIM"),
    ("gamma1", "method", "gamma1"),
    ("gamma2", "method", "gamma2"),
    ("alpha_n", "method", "alpha_n"),
    ("myLambda", "method", "myLambda"),
    ("max_iter", "method", "max_iter"),
    ("m1", "method", "m1"),
    ("m2", "method", "m2"),
    ("r", "experimental_settings", "r"),
    ("ch", "configs", "ch"),
)
_PER_IMAGE = ("PSNR", "SSIM", "PSNR_observation", "SSIM_observation")
_LAST_TITLE = "Result PSNR - Result SSIM - Observed PSNR - Observed SSIM (for each images)"


def get_csv_header() -> str:
    return ",".join(title for title, _, _ in _COLUMNS) + "," + _LAST_TITLE + "\n"


def get_csv_data(data: dict) -> str:
    cells = [str(data[section][key]) for _, section, key in _COLUMNS]
    for key in _PER_IMAGE:
        cells += [str(res[key]) for res in data["results"].values()]
    return "".join(c + "," for c in cells)


def get_csv_footer(data: dict) -> str:
    return "".join(str(res["filename"]) + "," for res in data["results"].values())


def touch_textfile(filepath) -> None:
    with open(filepath, "w") as f:
        f.write(get_csv_header())


def write_textfile(filepath, data: dict) -> None:
    with open(filepath, "a") as f:
        f.write(get_csv_data(data) + "\n")


def add_footer_textfile(filepath, data: dict) -> None:
    with open(filepath, "a") as f:
        f.write(get_csv_footer(data) + "\n")
