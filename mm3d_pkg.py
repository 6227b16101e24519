"""Imports the host-side binding that lives in ``map-merge_b200/`` (a hyphen is not importable by name)."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "map-merge_b200")


def load():
    if "map_merge_b200" in sys.modules:
        return sys.modules["map_merge_b200"]
    spec = importlib.util.spec_from_file_location("map_merge_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["map_merge_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_synth():
    load()
    import importlib
    return importlib.import_module("map_merge_b200.synth")
