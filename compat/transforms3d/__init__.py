"""Minimal transforms3d stand-in: the two matrix builders learning/spg.py:241-253 calls."""
from . import axangles, zooms  # noqa: F401
