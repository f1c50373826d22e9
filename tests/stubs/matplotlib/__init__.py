"""Minimal stand-in for `matplotlib` (absent from this image): opt/util/util.py imports pyplot at
module level and only uses it for the viridis colour map of depth images."""
