"""Import alias: the package directory is named ``improving-learned-index_b200`` (a hyphen is
not a valid module name), so ``import improving_learned_index_b200`` loads it from there."""
import importlib.util
import pathlib
import sys

_dir = pathlib.Path(__file__).resolve().with_name("improving-learned-index_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
