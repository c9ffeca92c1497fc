"""Imports the hyphen-named package directory `nerf-prv_b200/` as module `nerf_prv_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if "nerf_prv_b200" in sys.modules:
        return sys.modules["nerf_prv_b200"]
    pkg_dir = os.path.join(_ROOT, "nerf-prv_b200")
    spec = importlib.util.spec_from_file_location("nerf_prv_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["nerf_prv_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
