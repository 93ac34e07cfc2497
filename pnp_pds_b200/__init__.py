"""Import shim: the product package lives in ``pnp-pds_b200/`` (the name the build
contract fixes); a hyphen is not importable, so ``import pnp_pds_b200`` resolves its
submodules from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pnp-pds_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
