"""method -> (algorithm, denoiser) labels (utils/utils_method_master.py:4-21), both vocabularies."""
from ..engine import canonical_method

_TABLE = {
    "A-Proposed": ("PnP-PDS", "DnCNN"), "A-PnPFBS-DnCNN": ("PnP-FBS", "DnCNN"), "A-PnPPDS-BM3D": ("PnP-PDS", "BM3D"),
    "A-PnPFBS-BM3D": ("PnP-FBS", "BM3D"), "A-PDS-TV": ("PDS", ""), "A-RED-DnCNN": ("RED-SD", "DnCNN"),
    "A-PnPPDS-unstable-DnCNN": ("PnP-PDS", "DnCNN (unstable)"), "B-Proposed": ("PnP-PDS", "DnCNN"),
    "C-Proposed": ("PnP-PDS", "DnCNN"), "C-PnPPDS-BM3D": ("PnP-PDS", "BM3D"), "C-PnPADMM-DnCNN": ("PnP-ADMM", "DnCNN"),
    "C-RED-DnCNN": ("RED-ADMM", "DnCNN"), "C-PnP-unstable-DnCNN": ("PnP-PDS", "DnCNN (unstable)"),
}


def get_algorithm_denoiser(method):
    return _TABLE.get(canonical_method(method), ("unknown algorithm", "unknown denoiser"))
