"""Import alias: ``import vibo_b200`` -> the package in
``variational-item-response-theory-public_b200/`` (a directory name Python's
import statement cannot spell)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("variational-item-response-theory-public_b200")
sys.modules[__name__] = _pkg
