"""Importable alias for the product package.

The product lives in ``clifford-group-equivariant-simplicial-message-passing-networks_b200/``
(the directory name the build contract fixes; hyphens make it un-importable by name).
This alias points ``csmpn_b200.*`` at that directory, so
``from csmpn_b200.algebra.cliffordalgebra import CliffordAlgebra`` and
``from csmpn_b200.models.cegnn_utils import EGCL`` mirror the reference's
``csmpn.algebra`` / ``csmpn.models`` import paths.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "clifford-group-equivariant-simplicial-message-passing-networks_b200",
)
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
