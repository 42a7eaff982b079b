"""Import alias: the package directory is named `petsc-openacc_b200` (not a Python identifier)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_m = importlib.import_module("petsc-openacc_b200")
sys.modules[__name__] = _m
