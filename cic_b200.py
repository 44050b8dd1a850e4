"""Import shim: the package directory is `contextual-image-compression_b200` (not a Python identifier).

`import cic_b200` and `from cic_b200.gan import ...` resolve to the SAME module objects as the package's own names: every
sub-module is registered under both prefixes (without that, `from cic_b200.gan import x` would load a second copy of gan /
models / runtime with their own module-level state)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_name = "contextual-image-compression_b200"
_pkg = importlib.import_module(_name)
for _k, _m in list(sys.modules.items()):
    if _k.startswith(_name + "."):
        sys.modules[__name__ + _k[len(_name):]] = _m
sys.modules[__name__] = _pkg
