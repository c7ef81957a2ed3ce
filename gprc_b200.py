"""Alias: ``import gprc_b200`` == the package in ``gaussian-process-regression_b200/`` (its directory name is not a
valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("gaussian-process-regression_b200")
sys.modules[__name__] = _pkg
