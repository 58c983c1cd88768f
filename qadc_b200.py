"""Import alias: ``import qadc_b200`` == the package in ``quick-adc_b200/`` (whose directory
name keeps the reference's hyphen and therefore cannot be written in an import statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("quick-adc_b200")
