"""Import shim: the package directory `ekf-monoslam_for_3d-reconstruction_b200/` is not a valid
Python identifier, so it is registered here under the module name `ekf_b200`."""
import importlib.util
import os
import sys

_NAME = "ekf_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ekf-monoslam_for_3d-reconstruction_b200")


def load_package():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
