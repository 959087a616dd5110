"""
Makes the reference's `import gnn.mpnn` / `gnn.mpnn.GGNN(constants)` (Workflow.py:265-292)
resolve to the B200 drop-in classes without touching the reference tree:

    import graphinvent_b200.dropin as dropin; dropin.install()      # before `import Workflow`

The reference's S2V / AttentionS2V names are left undefined (they cannot be constructed in the
reference either, SURVEY.md §2 note a).
"""
import sys
import types

_NAMES = ("gnn", "gnn.mpnn", "gnn.modules")
_saved = {}


def install():
    from .gnn import modules, mpnn
    for n in _NAMES:
        _saved.setdefault(n, sys.modules.get(n))
    pkg = types.ModuleType("gnn")
    pkg.__path__ = []          # mark as package
    pkg.mpnn, pkg.modules = mpnn, modules
    sys.modules["gnn"] = pkg
    sys.modules["gnn.mpnn"] = mpnn
    sys.modules["gnn.modules"] = modules
    return pkg


def uninstall():
    for n in _NAMES:
        old = _saved.pop(n, None)
        if old is not None:
            sys.modules[n] = old
        else:
            sys.modules.pop(n, None)
