"""Environment producers are out of scope for this package ("stay unchanged").

``main.py`` of the reference does ``from sac.envs import *``.  The probe environments live in the reference
repository (sac/envs.py) and depend on gymnasium; when ``SAC_REFERENCE_ROOT`` points at a checkout of the
reference they are re-exported from there unmodified, otherwise importing a name raises ImportError."""
import importlib.util as _ilu
import os as _os

_root = _os.environ.get("SAC_REFERENCE_ROOT")
if _root and _os.path.exists(_os.path.join(_root, "sac", "envs.py")):
    _spec = _ilu.spec_from_file_location("_reference_sac_envs", _os.path.join(_root, "sac", "envs.py"))
    _mod = _ilu.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("_")})
    __all__ = [k for k in vars(_mod) if not k.startswith("_")]
else:
    __all__ = []
