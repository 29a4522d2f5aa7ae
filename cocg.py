"""Import shim: `import cocg` == the package in ./collaborative-circom_b200 (hyphenated directory name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("collaborative-circom_b200")
sys.modules[__name__] = _pkg
