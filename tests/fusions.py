"""The fusion definitions moved into the package (``fuser_b200/fusions.py``); the tests keep
importing them from here."""
from fuser_b200.fusions import *  # noqa: F401,F403
from fuser_b200.fusions import _down, _up  # noqa: F401
