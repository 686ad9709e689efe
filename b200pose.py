"""Import alias: ``import b200pose`` == the package ``3dhumanposeestimation_b200`` (whose name starts
with a digit and therefore cannot appear in a plain import statement)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("3dhumanposeestimation_b200")
