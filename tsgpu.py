"""Import shim: the package directory is named after the reference repo (hyphens), which `import` cannot spell."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("tiered-storage-for-apache-kafka_b200")
sys.modules[__name__] = _pkg
