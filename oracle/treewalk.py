"""Independent CPU restatement of the compiled part of the hot path (TEST INFRASTRUCTURE).

The reference's ``classifier.predict_proba(df[all_features])[:, 1]``
(``02-register-model.ipynb:335-337``) runs, inside un-vendored scikit-learn
(pinned 1.1.1 at ``app/requirements.txt:14``; algorithm read from the sklearn
installed in this image):

1. categorical ``SimpleImputer(constant "missing")`` then ``OneHotEncoder(handle_unknown="ignore")``
   (``01-train-model.ipynb:200-206``): unknown / missing category -> all-zero one-hot block;
2. numeric ``SimpleImputer(median)`` (``01-train-model.ipynb:212``): NaN -> training median;
3. hstack -> N x 85, cast to float32 (``sklearn/tree/_classes.py`` ``_validate_X_predict``);
4. per tree, root -> leaf with ``X[i, feature] <= threshold`` => left, X float32,
   threshold float64 (``sklearn/tree/_tree.pyx`` ``_apply_dense``);
5. RandomForest: float64 sum over trees of the leaf class fraction, then
   ``/= n_estimators`` (``sklearn/ensemble/_forest.py`` ``predict_proba``); label =
   argmax (class 1 iff p1 > p0);
   GBDT (BASELINE configs 2-4 only): ``raw = init + sum_t lr * value_t`` in tree
   order (``sklearn/ensemble/_gradient_boosting.pyx`` ``predict_stages``), proba =
   ``expit(raw)``, label = ``raw >= 0`` (``sklearn/ensemble/_gb.py`` ``predict``).

``dump_pipeline`` extracts plain arrays from a fitted pipeline; everything else
works on those arrays only (numpy here, C in ``oracle/c/forest_walk.c``).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import pandas as pd
from scipy.special import expit

from .reference_pipeline import CATEGORICAL_FEATURES, NUMERIC_FEATURES

RF_MEAN = 0
GBDT_LOGISTIC = 1


def dump_pipeline(pipeline) -> dict:
    """Fitted reference-style Pipeline -> plain arrays (oracle's own exchange format)."""
    pre = pipeline.named_steps["preprocessor"]
    clf = pipeline.named_steps["classifier"]
    ohe = pre.named_transformers_["categorical"].named_steps["ohe"]
    med = pre.named_transformers_["numeric"].named_steps["imputer"].statistics_
    cats = [np.asarray(c, dtype=object) for c in ohe.categories_]
    cat_sizes = np.array([len(c) for c in cats], dtype=np.int32)
    cat_offsets = np.concatenate([[0], np.cumsum(cat_sizes)]).astype(np.int32)

    if hasattr(clf, "estimators_") and clf.__class__.__name__ == "RandomForestClassifier":
        kind = RF_MEAN
        trees = [e.tree_ for e in clf.estimators_]
        init_raw, scale = 0.0, 1.0
        assert list(clf.classes_) == [0, 1]
    else:
        kind = GBDT_LOGISTIC
        trees = [e.tree_ for e in clf.estimators_[:, 0]]
        # init_ is a DummyClassifier(prior); raw init = logit(clip(p1)) -- _gb.py _init_raw_predictions
        p1 = float(clf.init_.class_prior_[1])
        eps = np.finfo(np.float64).eps
        p1 = min(max(p1, eps), 1 - eps)
        init_raw = float(np.log(p1 / (1 - p1)))
        scale = float(clf.learning_rate)

    left, right, feat, thr, val, off = [], [], [], [], [], [0]
    for t in trees:
        left.append(t.children_left.astype(np.int32))
        right.append(t.children_right.astype(np.int32))
        feat.append(t.feature.astype(np.int32))
        thr.append(t.threshold.astype(np.float64))
        if kind == RF_MEAN:
            v = t.value[:, 0, :]  # counts (sklearn<1.4) or fractions (>=1.4): normalise either way
            s = v.sum(axis=1)
            s[s == 0.0] = 1.0
            val.append((v[:, 1] / s).astype(np.float64))
        else:
            val.append(t.value[:, 0, 0].astype(np.float64))
        off.append(off[-1] + t.node_count)
    return dict(
        kind=kind,
        n_trees=len(trees),
        init_raw=init_raw,
        scale=scale,
        categories=cats,
        cat_sizes=cat_sizes,
        cat_offsets=cat_offsets,
        medians=np.asarray(med, dtype=np.float64),
        tree_off=np.asarray(off, dtype=np.int64),
        left=np.concatenate(left),
        right=np.concatenate(right),
        feature=np.concatenate(feat),
        threshold=np.concatenate(thr),
        value=np.concatenate(val),
    )


def encode_frame(dump: dict, df: pd.DataFrame):
    """DataFrame (any column order) -> (codes int32 N x 9 with -1 = unknown/missing,
    nums float64 N x 14 with NaN = missing).  Restates steps 1-2's *lookup* only."""
    n = len(df)
    codes = np.full((n, len(CATEGORICAL_FEATURES)), -1, dtype=np.int32)
    for j, name in enumerate(CATEGORICAL_FEATURES):
        cats = dump["categories"][j]
        lut = {c: i for i, c in enumerate(cats.tolist())}
        col = df[name].tolist()
        codes[:, j] = [lut.get(v, -1) if isinstance(v, str) else -1 for v in col]
    nums = np.empty((n, len(NUMERIC_FEATURES)), dtype=np.float64)
    for j, name in enumerate(NUMERIC_FEATURES):
        nums[:, j] = pd.to_numeric(df[name]).to_numpy(dtype=np.float64)
    return codes, nums


def transform_dense(dump: dict, codes: np.ndarray, nums: np.ndarray) -> np.ndarray:
    """Steps 1-3: -> dense float32 N x 85 exactly as the forest sees it."""
    n = codes.shape[0]
    n_ohe = int(dump["cat_offsets"][-1])
    X = np.zeros((n, n_ohe + nums.shape[1]), dtype=np.float64)
    rows = np.arange(n)
    for j in range(codes.shape[1]):
        ok = codes[:, j] >= 0
        X[rows[ok], dump["cat_offsets"][j] + codes[ok, j]] = 1.0
    filled = np.where(np.isnan(nums), dump["medians"][None, :], nums)
    X[:, n_ohe:] = filled
    with np.errstate(over="ignore"):
        X32 = X.astype(np.float32)
    if not np.isfinite(X32).all():
        # sklearn: "Input X contains infinity or a value too large for dtype('float32')."
        raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
    return X32


def walk_numpy(dump: dict, X32: np.ndarray):
    """Steps 4-5 in numpy.  Returns (proba1 float64, label int32, raw float64)."""
    n = X32.shape[0]
    acc = np.zeros(n, dtype=np.float64)
    if dump["kind"] == GBDT_LOGISTIC:
        acc += dump["init_raw"]
    rows = np.arange(n)
    for t in range(dump["n_trees"]):
        lo = int(dump["tree_off"][t])
        L = dump["left"][lo:]
        R = dump["right"][lo:]
        F = dump["feature"][lo:]
        T = dump["threshold"][lo:]
        V = dump["value"][lo:]
        node = np.zeros(n, dtype=np.int64)
        while True:
            is_leaf = L[node] == -1
            if is_leaf.all():
                break
            f = np.where(is_leaf, 0, F[node])
            go_left = X32[rows, f].astype(np.float64) <= T[node]
            nxt = np.where(go_left, L[node], R[node])
            node = np.where(is_leaf, node, nxt)
        if dump["kind"] == RF_MEAN:
            acc += V[node]
        else:
            acc += dump["scale"] * V[node]
    if dump["kind"] == RF_MEAN:
        s1 = acc
        s0 = float(dump["n_trees"]) - s1
        proba1 = s1 / float(dump["n_trees"])
        label = (s1 > s0).astype(np.int32)
        return proba1, label, acc
    proba1 = expit(acc)
    label = (acc >= 0).astype(np.int32)
    return proba1, label, acc


def predict_numpy(dump: dict, df: pd.DataFrame):
    codes, nums = encode_frame(dump, df)
    return walk_numpy(dump, transform_dense(dump, codes, nums))[:2]


# --------------------------------------------------------------------------
# C restatement (oracle/c/forest_walk.c) -- multi-threaded CPU baseline.
# --------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_forest.so")
_lib = None


def build_c(force: bool = False) -> str:
    """Compile the C restatement (gcc -O3 -fopenmp).  Building the checker is not using it."""
    src = os.path.join(_HERE, "c", "forest_walk.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"]
        )
    return _LIB_PATH


def _load_c():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build_c()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_forest_predict.restype = ctypes.c_int
    return _lib


def predict_c(dump: dict, codes: np.ndarray, nums: np.ndarray, threads: int = 0):
    """C restatement of steps 1-5 on encoded rows.  Returns (proba1 f64, label i32)."""
    lib = _load_c()
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    nums = np.ascontiguousarray(nums, dtype=np.float64)
    n = codes.shape[0]
    proba = np.empty(n, dtype=np.float64)
    label = np.empty(n, dtype=np.int32)
    arrs = dict(
        cat_offsets=np.ascontiguousarray(dump["cat_offsets"], dtype=np.int32),
        medians=np.ascontiguousarray(dump["medians"], dtype=np.float64),
        tree_off=np.ascontiguousarray(dump["tree_off"], dtype=np.int64),
        left=np.ascontiguousarray(dump["left"], dtype=np.int32),
        right=np.ascontiguousarray(dump["right"], dtype=np.int32),
        feature=np.ascontiguousarray(dump["feature"], dtype=np.int32),
        threshold=np.ascontiguousarray(dump["threshold"], dtype=np.float64),
        value=np.ascontiguousarray(dump["value"], dtype=np.float64),
    )
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.oracle_forest_predict(
        ctypes.c_int(int(dump["kind"])),
        ctypes.c_int(int(dump["n_trees"])),
        ctypes.c_double(float(dump["init_raw"])),
        ctypes.c_double(float(dump["scale"])),
        ctypes.c_int(codes.shape[1]),
        ctypes.c_int(nums.shape[1]),
        p(arrs["cat_offsets"]),
        p(arrs["medians"]),
        p(arrs["tree_off"]),
        p(arrs["left"]),
        p(arrs["right"]),
        p(arrs["feature"]),
        p(arrs["threshold"]),
        p(arrs["value"]),
        p(codes),
        p(nums),
        ctypes.c_int64(n),
        p(proba),
        p(label),
        ctypes.c_int(int(threads)),
    )
    if rc == -2:
        raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
    if rc != 0:
        raise RuntimeError(f"oracle_forest_predict failed rc={rc}")
    return proba, label
