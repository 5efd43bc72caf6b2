"""Import alias: ``import use_b200`` loads the package in ``universal-speech-enhancement_b200/`` (a directory name
Python cannot import directly)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "universal-speech-enhancement_b200")
_spec = importlib.util.spec_from_file_location("use_b200", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["use_b200"] = _mod
_spec.loader.exec_module(_mod)
