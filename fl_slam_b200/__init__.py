"""Importable alias of the package directory ``fl-slam_b200/`` (a hyphen cannot appear in a Python module name).

``import fl_slam_b200`` executes ``fl-slam_b200/__init__.py`` with this module as its namespace and points the
package search path at that directory, so ``fl_slam_b200.operators`` etc. resolve to the files there.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fl-slam_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _os, _f
