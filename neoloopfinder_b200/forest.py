"""Host-side flattening of the Peakachu random forest into the GPU node array.

Replaces ``Peakachu.load_models`` (neoloop/callers.py:572-577), which unpickles the model for
every assembly, by one flattening per model file and device: the sklearn estimators are turned
into the concatenated arrays ``nlf_forest_create`` takes (include/nlf_b200.h) and live on the
GPU behind a handle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

__all__ = ["flatten_forest", "DeviceForest", "load_forest"]


def flatten_forest(model):
    """sklearn ForestClassifier -> dict of concatenated per-node arrays.

    Semantics restated from sklearn (1.9): ``predict_proba`` casts X to float32, each tree's
    ``_apply_dense`` compares the float32 feature with the float64 threshold (``<=`` goes left,
    NaN follows ``missing_go_to_left``), the leaf's ``value[leaf, 0, :]`` (already a class
    fraction) is accumulated over the estimators in order and divided by their number
    (sklearn/ensemble/_forest.py predict_proba / _accumulate_prediction,
    sklearn/tree/_tree.pyx _apply_dense).
    """
    if not hasattr(model, "estimators_"):
        raise TypeError("the model must be a fitted sklearn forest classifier (estimators_ missing)")
    n_jobs = getattr(model, "n_jobs", None)
    classes = list(getattr(model, "classes_", [0, 1]))
    if len(classes) != 2:
        raise ValueError(f"binary loop/non-loop forest expected, got classes {classes}")
    k1 = 1  # predict_proba(...)[:, 1]
    off = [0]
    left, right, feat, thr, ml, p1 = [], [], [], [], [], []
    for est in model.estimators_:
        t = est.tree_
        left.append(np.asarray(t.children_left, dtype=np.int32))
        right.append(np.asarray(t.children_right, dtype=np.int32))
        feat.append(np.asarray(t.feature, dtype=np.int32))
        thr.append(np.asarray(t.threshold, dtype=np.float64))
        mgl = getattr(t, "missing_go_to_left", None)
        ml.append(np.zeros(t.node_count, np.uint8) if mgl is None else np.asarray(mgl, dtype=np.uint8))
        p1.append(np.asarray(t.value[:, 0, k1], dtype=np.float64))
        off.append(off[-1] + int(t.node_count))
    n_features = int(model.n_features_in_) if hasattr(model, "n_features_in_") else int(model.feature_importances_.size)
    return {
        "n_trees": len(model.estimators_), "n_features": n_features,
        "tree_off": np.asarray(off, dtype=np.int64),
        "left": np.concatenate(left), "right": np.concatenate(right), "feature": np.concatenate(feat),
        "threshold": np.concatenate(thr), "missing_left": np.concatenate(ml), "leaf_p1": np.concatenate(p1),
        "order_deterministic": n_jobs in (None, 1),
    }


class DeviceForest:
    """Forest resident on one GPU."""

    def __init__(self, flat: dict, ctx: "_lib.Context" = None):
        self.ctx = ctx or _lib.Context.get()
        self.flat = flat
        self.n_features = flat["n_features"]
        self.n_trees = flat["n_trees"]
        self.w = int((np.sqrt(self.n_features) - 1) / 2)      # callers.py:577
        self.h = C.c_void_p()
        L = _lib.load()
        _lib.check(L.nlf_forest_create(
            self.ctx.h, flat["n_trees"], _lib.ptr(flat["tree_off"], C.c_int64), _lib.ptr(flat["left"], C.c_int),
            _lib.ptr(flat["right"], C.c_int), _lib.ptr(flat["feature"], C.c_int),
            _lib.ptr(flat["threshold"], C.c_double), _lib.ptr(flat["missing_left"], C.c_ubyte),
            _lib.ptr(flat["leaf_p1"], C.c_double), flat["n_features"], C.byref(self.h)))

    def predict_proba1(self, X) -> np.ndarray:
        """``model.predict_proba(X)[:, 1]`` on the GPU (X is cast to float32 like sklearn does)."""
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.ndim != 2 or X.shape[1] != self.n_features:
            raise ValueError(f"X has shape {X.shape}, expected (*, {self.n_features})")
        if np.isinf(X).any():
            raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
        out = np.empty(X.shape[0], dtype=np.float64)
        _lib.check(_lib.load().nlf_forest_predict(self.h, _lib.ptr(X, C.c_float), X.shape[0], _lib.ptr(out, C.c_double)))
        return out

    def close(self):
        if self.h:
            _lib.load().nlf_forest_destroy(self.h)
            self.h = C.c_void_p()


_cache = {}


def load_forest(model_or_path, ctx=None) -> DeviceForest:
    """Flatten once per (model file, device, process); later calls return the resident forest."""
    ctx = ctx or _lib.Context.get()
    if isinstance(model_or_path, (str, os.PathLike)):
        path = os.path.abspath(os.fspath(model_or_path))
        key = (os.getpid(), ctx.device, path, os.path.getmtime(path))
        if key not in _cache:
            import joblib
            _cache[key] = DeviceForest(flatten_forest(joblib.load(path)), ctx)
        return _cache[key]
    key = (os.getpid(), ctx.device, id(model_or_path))
    if key not in _cache:
        df = DeviceForest(flatten_forest(model_or_path), ctx)
        df._model_ref = model_or_path       # keep id() stable
        _cache[key] = df
    return _cache[key]
