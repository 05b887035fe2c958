"""Import helper: the package directory is named `dav1d-mirror_b200` (not a
valid Python identifier), so it is registered under the importable alias
`dav1d_mirror_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
_ALIAS = "dav1d_mirror_b200"


def load_pkg():
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    path = os.path.join(ROOT, "dav1d-mirror_b200")
    spec = importlib.util.spec_from_file_location(
        _ALIAS, os.path.join(path, "__init__.py"), submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
