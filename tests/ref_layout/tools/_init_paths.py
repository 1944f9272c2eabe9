"""Stand-in for the reference's tools/_init_paths.py: `lib/` goes to sys.path[0] (which is why shadowing the reference's
modules by PYTHONPATH cannot work and the shim rebinds names instead)."""
import os
import sys

_lib = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lib"))
if _lib not in sys.path:
    sys.path.insert(0, _lib)
