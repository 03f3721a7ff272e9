"""Import alias: the package directory is named `game-ai_b200` (not a Python identifier)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("game-ai_b200")
sys.modules[__name__] = _pkg
