"""Imports the package directory `singular-demux_b200/` (not a valid Python identifier) under the module
name `singular_demux_b200`."""
import importlib.util
import os
import sys

NAME = "singular_demux_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "singular-demux_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(_PKG, "__init__.py"),
                                                  submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
