"""NumPy-backed stand-in for the tiny slice of the ``jax`` API that the reference's hot-path
modules touch.  TEST INFRASTRUCTURE ONLY.

Why this exists: the reference (whabacivch/FL-SLAM) is Python + JAX, JAX is not installable in
the build container (no network), so the reference's own operator source cannot execute as-is.
Installing this shim as ``sys.modules['jax']`` lets ``tests/golden/generate_goldens.py`` import the
reference's *unmodified* source files from ``/root/reference`` and execute them with NumPy
(float64, same LAPACK family: gesdd / syevd / getrf) as the array backend.  The numbers it
produces are committed under ``tests/golden/*.npz`` and pin ``oracle/`` (see the header there).

Nothing in the product (``fl-slam_b200/``), in ``bench.py`` or in the ``-m gpu`` tests imports
this file; it never travels into a timed or shipped path.

Semantics implemented (and only these):
  * ``jax.jit`` / ``jax.jit(static_argnames=...)``   -> identity decorator
  * ``jax.vmap(f, in_axes)``                          -> Python loop + ``np.stack`` over pytrees
  * ``jax.nn.softmax`` (max-shifted), ``jax.nn.sigmoid`` (logistic), ``jax.nn.softplus``
  * ``jax.lax.scan``                                  -> Python loop
  * ``jax.scipy.linalg.solve_triangular``, ``jax.scipy.special.logsumexp``
  * ``jax.numpy``: NumPy functions whose ndarray results are re-viewed as ``Arr`` so that the
    functional update syntax ``x.at[idx].set(v)`` / ``.add(v)`` works and returns a copy.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import scipy.linalg
import scipy.special


class _AtIndexer:
    __slots__ = ("_arr", "_idx")

    def __init__(self, arr, idx):
        self._arr = arr
        self._idx = idx

    def set(self, value):
        out = np.array(self._arr, copy=True).view(Arr)
        out[self._idx] = value
        return out

    def add(self, value):
        out = np.array(self._arr, copy=True).view(Arr)
        out[self._idx] = out[self._idx] + value
        return out


class _At:
    __slots__ = ("_arr",)

    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIndexer(self._arr, idx)


class Arr(np.ndarray):
    """ndarray with a JAX-style immutable-update ``.at`` property."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, tuple):
        if hasattr(x, "_fields"):  # namedtuple (e.g. SVDResult / EighResult): plain tuple is enough
            return tuple(_wrap(v) for v in x)
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    if isinstance(x, np.generic):
        # JAX returns 0-d arrays for scalar reductions; callers do float(x) / int(x) on them.
        return np.asarray(x).view(Arr)
    return x


def _wrapping(fn):
    def inner(*a, **k):
        return _wrap(fn(*a, **k))

    inner.__name__ = getattr(fn, "__name__", "wrapped")
    return inner


class _WrappedModule(types.ModuleType):
    def __init__(self, name, backing):
        super().__init__(name)
        self._backing = backing

    def __getattr__(self, item):
        v = getattr(self._backing, item)
        if isinstance(v, type) or not callable(v):
            return v
        return _wrapping(v)


def _jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def _tree_map(fn, tree):
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(fn, t) for t in tree) if not hasattr(tree, "_fields") else tuple(
            _tree_map(fn, t) for t in tree
        )
    return fn(tree)


def _tree_stack(items):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return tuple(_tree_stack([it[i] for it in items]) for i in range(len(first)))
    return _wrap(np.stack([np.asarray(it) for it in items], axis=0))


def _vmap(fn, in_axes=0, out_axes=0):
    assert out_axes == 0

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                assert ax == 0
                n = np.asarray(a).shape[0]
                break
        outs = []
        for i in range(n):
            call = [(_wrap(np.asarray(a)[i]) if ax is not None else a) for a, ax in zip(args, axes)]
            outs.append(fn(*call))
        return _tree_stack(outs)

    return mapped


def _softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float64)
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return _wrap(e / np.sum(e, axis=axis, keepdims=True))


def _sigmoid(x):
    return _wrap(np.asarray(scipy.special.expit(np.asarray(x, dtype=np.float64))))


def _softplus(x):
    return _wrap(np.logaddexp(np.asarray(x, dtype=np.float64), 0.0))


def _scan(f, init, xs, length=None):
    carry = init
    ys = []
    if xs is None:
        n = length
        get = lambda i: None
    else:
        leaves = xs if isinstance(xs, (tuple, list)) else (xs,)
        n = np.asarray(leaves[0]).shape[0]
        if isinstance(xs, (tuple, list)):
            get = lambda i: tuple(_wrap(np.asarray(v)[i]) for v in xs)
        else:
            get = lambda i: _wrap(np.asarray(xs)[i])
    for i in range(n):
        carry, y = f(carry, get(i))
        ys.append(y)
    if ys and ys[0] is not None:
        ys = _tree_stack(ys)
    else:
        ys = None
    return carry, ys


def install():
    """Register the shim as ``jax`` / ``jax.numpy`` / … in ``sys.modules`` (idempotent)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "__is_numpy_shim__", False):
        return sys.modules["jax"]
    if "jax" in sys.modules:
        raise RuntimeError("a real jax is already imported; the shim must not shadow it")

    jnp = _WrappedModule("jax.numpy", np)
    jnp.ndarray = np.ndarray
    jnp.linalg = _WrappedModule("jax.numpy.linalg", np.linalg)
    jnp.float64, jnp.float32, jnp.int32, jnp.int64, jnp.uint8, jnp.bool_ = (
        np.float64, np.float32, np.int32, np.int64, np.uint8, np.bool_)
    jnp.pi, jnp.inf, jnp.nan, jnp.newaxis = np.pi, np.inf, np.nan, np.newaxis
    jnp.finfo, jnp.iinfo = np.finfo, np.iinfo

    def _array(x, dtype=None, **kw):
        return _wrap(np.array(x, dtype=dtype))

    def _asarray(x, dtype=None, **kw):
        return _wrap(np.asarray(x, dtype=dtype))

    jnp.array = _array
    jnp.asarray = _asarray

    jax = types.ModuleType("jax")
    jax.__is_numpy_shim__ = True
    jax.numpy = jnp
    jax.jit = _jit
    jax.vmap = _vmap
    jax.Array = np.ndarray

    nn = types.ModuleType("jax.nn")
    nn.softmax, nn.sigmoid, nn.softplus = _softmax, _sigmoid, _softplus
    jax.nn = nn

    lax = types.ModuleType("jax.lax")
    lax.scan = _scan
    lax.stop_gradient = lambda x: x
    jax.lax = lax

    jsp = types.ModuleType("jax.scipy")
    jsp_linalg = types.ModuleType("jax.scipy.linalg")
    jsp_linalg.solve_triangular = _wrapping(
        lambda a, b, lower=False, **kw: scipy.linalg.solve_triangular(np.asarray(a), np.asarray(b), lower=lower))
    jsp_special = types.ModuleType("jax.scipy.special")
    jsp_special.logsumexp = _wrapping(scipy.special.logsumexp)
    jsp.linalg, jsp.special = jsp_linalg, jsp_special
    jax.scipy = jsp

    class _Config:
        def update(self, *a, **k):
            return None

    jax.config = _Config()
    jax.devices = lambda *a, **k: ["numpy-shim"]
    jax.tree_util = types.ModuleType("jax.tree_util")
    jax.tree_util.tree_map = _tree_map
    jax.block_until_ready = lambda x: x

    for name, mod in [("jax", jax), ("jax.numpy", jnp), ("jax.numpy.linalg", jnp.linalg), ("jax.nn", nn),
                      ("jax.lax", lax), ("jax.scipy", jsp), ("jax.scipy.linalg", jsp_linalg),
                      ("jax.scipy.special", jsp_special), ("jax.tree_util", jax.tree_util)]:
        sys.modules[name] = mod
    return jax


def load_reference(reference_root="/root/reference", n_bins=48, tau=0.1):
    """Import the reference's hot-path modules (unmodified, from disk) on top of the shim.

    ``fl_slam_poc`` and its sub-packages are registered as bare namespace modules so their
    ``__init__.py`` files (which pull in ROS-facing code) do not run; only the leaf modules the
    hot path needs are executed.  The two constants the archived modules expect but the tree no
    longer defines are injected: ``GC_B_BINS`` and ``GC_TAU_SOFT_ASSIGN``.
    Returns a dict of the loaded modules.
    """
    import importlib
    import importlib.util
    import os

    install()
    pkg_root = os.path.join(reference_root, "fl_ws/src/fl_slam_poc/fl_slam_poc")
    for name, rel in [("fl_slam_poc", ""), ("fl_slam_poc.common", "common"),
                      ("fl_slam_poc.common.geometry", "common/geometry"),
                      ("fl_slam_poc.backend", "backend"), ("fl_slam_poc.backend.operators", "backend/operators"),
                      ("fl_slam_poc.backend.structures", "backend/structures")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(pkg_root, rel)]
            sys.modules[name] = m
    constants = importlib.import_module("fl_slam_poc.common.constants")
    constants.GC_B_BINS = int(n_bins)
    constants.GC_TAU_SOFT_ASSIGN = float(tau)
    sys.modules["fl_slam_poc.common"].constants = constants
    mods = {"constants": constants}
    for short, full in [("certificates", "fl_slam_poc.common.certificates"),
                        ("primitives", "fl_slam_poc.common.primitives"),
                        ("se3_jax", "fl_slam_poc.common.geometry.se3_jax"),
                        ("belief", "fl_slam_poc.common.belief"),
                        ("kappa", "fl_slam_poc.backend.operators.kappa"),
                        ("imu_preintegration", "fl_slam_poc.backend.operators.imu_preintegration"),
                        ("point_budget", "fl_slam_poc.backend.operators.point_budget"),
                        ("deskew_constant_twist", "fl_slam_poc.backend.operators.deskew_constant_twist"),
                        ("mn_iw_struct", "fl_slam_poc.backend.structures.measurement_noise_iw_jax"),
                        ("mn_iw", "fl_slam_poc.backend.operators.measurement_noise_iw_jax")]:
        mods[short] = importlib.import_module(full)
        parent, leaf = full.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, mods[short])

    def _load_by_path(modname, path):
        spec = importlib.util.spec_from_file_location(modname, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m

    arch = os.path.join(reference_root, "archive")
    mods["binning"] = _load_by_path("ref_archive_binning", os.path.join(arch, "legacy_operators/binning.py"))
    mods["bin_atlas"] = _load_by_path("ref_archive_bin_atlas", os.path.join(arch, "bin_atlas.py"))
    mods["matrix_fisher_evidence"] = _load_by_path(
        "ref_archive_matrix_fisher_evidence", os.path.join(arch, "legacy_operators/matrix_fisher_evidence.py"))
    return mods
