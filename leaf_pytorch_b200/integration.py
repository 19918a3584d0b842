"""Wire the B200 frontend into an importable checkout of SarthakYadav/leaf-pytorch without editing it.

``install()`` rebinds the two names through which the reference reaches its frontend:
``models.classifier.get_frontend`` (used at reference models/classifier.py:11) and
``leaf_pytorch.get_frontend`` / ``leaf_pytorch.frontend.Leaf`` (reference leaf_pytorch/__init__.py:1,
frontend.py:22).  ``uninstall()`` restores them.  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, Tuple

from .frontend import Leaf
from .frontend_helper import get_frontend

_saved: Dict[Tuple[str, str], object] = {}


def _rebind(module_name: str, attr: str, value) -> bool:
    try:
        mod = sys.modules.get(module_name) or importlib.import_module(module_name)
    except Exception:                                    # module not importable in this environment
        return False
    if not hasattr(mod, attr):
        return False
    _saved.setdefault((module_name, attr), getattr(mod, attr))
    setattr(mod, attr, value)
    return True


def install(algo: str = "auto") -> Dict[str, bool]:
    """Patch the reference (must be on sys.path).  Returns which bindings were replaced."""
    def factory(opt):
        return get_frontend(opt, algo=algo)
    return {
        "models.classifier.get_frontend": _rebind("models.classifier", "get_frontend", factory),
        "leaf_pytorch.get_frontend": _rebind("leaf_pytorch", "get_frontend", factory),
        "leaf_pytorch.frontend_helper.get_frontend": _rebind("leaf_pytorch.frontend_helper", "get_frontend", factory),
        "leaf_pytorch.frontend.Leaf": _rebind("leaf_pytorch.frontend", "Leaf", Leaf),
    }


def uninstall() -> None:
    for (module_name, attr), value in list(_saved.items()):
        mod = sys.modules.get(module_name)
        if mod is not None:
            setattr(mod, attr, value)
        del _saved[(module_name, attr)]
