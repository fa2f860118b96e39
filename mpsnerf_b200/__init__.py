"""Importable alias for the package directory ``mps-nerf_b200/``.

The package directory name required by the repo layout contains a hyphen and
so cannot be imported directly; this shim points the import system at it.
"""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "mps-nerf_b200"
__path__.insert(0, str(_real))
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
