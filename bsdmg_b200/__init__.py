"""Importable alias of the ``bevy-signed-distance-mesh-generation_b200`` package directory (its name, which follows
the reference repository's, is not a valid Python identifier)."""
import importlib.util
import pathlib
import sys

_pkg_dir = pathlib.Path(__file__).resolve().parent.parent / "bevy-signed-distance-mesh-generation_b200"
_spec = importlib.util.spec_from_file_location(__name__, _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
