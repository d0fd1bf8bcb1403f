"""Import alias: the package directory `2d_multigrid_b200` is not a valid Python identifier, so
`import mg2d` (or importlib.import_module("2d_multigrid_b200")) is how user code reaches it."""
import importlib
import sys

_pkg = importlib.import_module("2d_multigrid_b200")
sys.modules[__name__] = _pkg


if __name__ == "__main__":          # `python mg2d.py ...` == `python -m 2d_multigrid_b200 ...`
    import runpy
    runpy.run_module("2d_multigrid_b200", run_name="__main__")
