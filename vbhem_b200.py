"""Import alias: ``import vbhem_b200`` loads the package that lives in the directory
``clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200/``
(a name Python cannot import directly because of the hyphens)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                     "clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200")
_spec = _u.spec_from_file_location("vbhem_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["vbhem_b200"] = _mod
_spec.loader.exec_module(_mod)
