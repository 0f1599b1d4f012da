"""Importable alias: the sources live in ``finetune-fair-diffusion_b200/`` (a directory name that
is not a Python identifier).  ``import fairguide`` loads that directory as this package."""
import importlib.util
import os
import sys

_SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finetune-fair-diffusion_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_SRC, "__init__.py"), submodule_search_locations=[_SRC])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
