"""Import helper: the product package lives in the directory ``tulip.jl_b200/`` (name fixed by the
project layout); a dot is not legal in a Python package name, so this registers it as
``tulip_jl_b200``.  Usage::

    import tlpb200_loader; pkg = tlpb200_loader.load()      # or: import tulip_jl_b200 afterwards
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "tulip.jl_b200")
NAME = "tulip_jl_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(NAME, None)
        raise
    return mod
