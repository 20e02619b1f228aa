"""Importable alias for the package directory `improving-inductive-oov-recsys_b200/` (whose name is
not a Python identifier): `import oov_b200` gives that package and registers every submodule under
the alias too, so `from oov_b200.inductive.get_inductive import get_inductive_embedder` works."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_REAL = "improving-inductive-oov-recsys_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
