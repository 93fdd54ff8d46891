"""Importable alias of the package directory `gan-rl_3d_b200/` (a hyphen is not a valid identifier):
`import gan_rl_3d_b200 as rlg` gives the package itself."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("gan-rl_3d_b200")
sys.modules[__name__] = _pkg
