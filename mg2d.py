"""Import alias: the package directory `2d_multigrid_b200` is not a valid Python identifier, so
`import mg2d` (or importlib.import_module("2d_multigrid_b200")) is how user code reaches it."""
import importlib
import sys

_pkg = importlib.import_module("2d_multigrid_b200")
sys.modules[__name__] = _pkg
