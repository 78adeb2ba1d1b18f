"""vbhem_b200 -- B200-native (sm_100a CUDA) E-step of VBHEM-H3M behind a C ABI.

Import name: ``vbhem_b200`` (see the alias module ``vbhem_b200.py`` at the repo root; the
package directory carries the long hyphenated project name).
"""
from . import pack, synth  # noqa: F401  (host-side layouts and synthetic data; no compute)

__all__ = ["pack", "synth", "api", "BaseSet"]


def __getattr__(name):
    if name in ("api", "BaseSet"):
        import importlib
        api = importlib.import_module(".api", __name__)
        return api if name == "api" else api.BaseSet
    raise AttributeError(name)
