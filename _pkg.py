"""Loader for the product package.

The package directory is named ``quadruped-robot_b200`` (the name the project brief prescribes); a
hyphen is not importable, so it is registered in ``sys.modules`` as ``quadruped_robot_b200``.
"""
import importlib.util
import os
import sys

_NAME = "quadruped_robot_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quadruped-robot_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
