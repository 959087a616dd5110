"""Drop-in replacement of the reference's `gnn` package for the hot path (see mpnn.py)."""
from . import modules, mpnn  # noqa: F401
