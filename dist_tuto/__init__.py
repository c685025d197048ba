"""Import shim: makes ``import dist_tuto.pth_b200`` resolve to the in-tree
package directory literally named ``dist_tuto.pth_b200/`` (a dotted directory
name is not importable by the normal path finder).

The real code lives in ``<repo>/dist_tuto.pth_b200/``; this module only
registers it in ``sys.modules`` under the dotted name so that pickled
references (``torch.multiprocessing`` spawn) and ``python -m`` both work.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_NAME = "dist_tuto.pth_b200"
_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), _PKG_NAME)


def _load():
    if _PKG_NAME in _sys.modules:
        return _sys.modules[_PKG_NAME]
    init = _os.path.join(_PKG_DIR, "__init__.py")
    if not _os.path.isfile(init):  # pragma: no cover - broken checkout
        raise ImportError(f"{_PKG_NAME}: package directory not found at {_PKG_DIR}")
    spec = _ilu.spec_from_file_location(_PKG_NAME, init, submodule_search_locations=[_PKG_DIR])
    mod = _ilu.module_from_spec(spec)
    _sys.modules[_PKG_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        _sys.modules.pop(_PKG_NAME, None)
        raise
    return mod


pth_b200 = _load()
