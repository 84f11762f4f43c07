"""Importable alias of the package directory `text-based-image-style-transfer_b200/` (whose name is not a
valid Python identifier): `import nst_b200` returns that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("text-based-image-style-transfer_b200")
