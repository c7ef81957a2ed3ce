"""Alias: ``import gprc_b200`` == the package in ``gaussian-process-regression_b200/`` (its directory name is not a
valid Python identifier).

The alias covers the submodules too: ``gprc_b200.fit`` must be THE module object ``gaussian-process-regression_b200.fit``.
Without that, ``from gprc_b200.fit import X`` executed ``fit.py`` a second time under the alias name and the import
machinery then rebound the package attribute ``fit`` from the function ``fit()`` to that second module (round 1: this
is what turned the GPU tier red).  Already-imported submodules are registered in ``sys.modules`` here; a meta-path
finder resolves the ones that are imported later (``gprc_b200.dist``, ``gprc_b200._quad`` ...) to the real module."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_REAL = "gaussian-process-regression_b200"
_ALIAS = __name__

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_REAL)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real):
        self._real = real

    def create_module(self, spec):
        return importlib.import_module(self._real)

    def exec_module(self, module):  # the real module is already executed
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            importlib.import_module(real)
        except ModuleNotFoundError:
            return None
        return importlib.machinery.ModuleSpec(fullname, _AliasLoader(real))


for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + ".") and _mod is not None:
        sys.modules[_ALIAS + _name[len(_REAL):]] = _mod
if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[_ALIAS] = _pkg
