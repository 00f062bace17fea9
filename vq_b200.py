"""Import shim: the package directory is named `multi-source-lms-for-audio_b200` (not a valid identifier), so
`import vq_b200` gives the same module object under an importable name."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("multi-source-lms-for-audio_b200")
sys.modules[__name__] = _pkg
