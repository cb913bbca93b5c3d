"""Import alias: the package directory is named ``tfe-slam-ucl_b200`` (not a valid Python identifier), so
``import tfe_slam_ucl_b200`` loads ``tfe-slam-ucl_b200/__init__.py`` under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tfe-slam-ucl_b200")
_spec = importlib.util.spec_from_file_location("tfe_slam_ucl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tfe_slam_ucl_b200"] = _mod
_spec.loader.exec_module(_mod)
