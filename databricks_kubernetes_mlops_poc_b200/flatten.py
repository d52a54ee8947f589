"""Fitted sklearn Pipeline -> forest blob (the layout in ``csrc/forest_blob.h``).

The reference serves ``artifacts/classifier/model/model.pkl`` -- a pickled sklearn
``Pipeline(ColumnTransformer -> RandomForestClassifier)`` (definition: reference
``databricks/src/01-train-model.ipynb:195-231``; loaded at
``02-register-model.ipynb:317-321``).  This module turns such a fitted object into
the flat, versioned binary the GPU engine consumes; it is the only place that knows
sklearn's attribute names.  It does arithmetic-free bookkeeping only:

* one-hot column ``j`` of the ColumnTransformer output is rewritten as
  ``(categorical row word f, category code c)``: the split ``x_j <= 0.5`` becomes
  "second child iff code[f] == c"  (unknown / missing category = code -1 = all-zero
  one-hot block = always first child, which is ``handle_unknown="ignore"``,
  ``01-train-model.ipynb:203-206``);
* float64 thresholds become the float32 ``t' = nextup(floor32(t64))``: for float32 ``x``,
  ``x <= t64  <=>  x <= max{f32 <= t64}  <=>  x < t'`` (sklearn compares float32 X with
  float64 thresholds, ``sklearn/tree/_tree.pyx`` ``_apply_dense``), so the kernel's test is
  "second child iff ``x >= t'``";
* leaf payloads stay float64: RF class-1 fraction, or ``learning_rate * value`` for GBDT
  (the product sklearn's ``predict_stages`` forms before adding);
* nodes are re-numbered breadth-first so siblings are adjacent, leaves become
  self-looping slots, and 32 trees are interleaved per group (see ``forest_blob.h``).
"""

from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field

import numpy as np

ROW_WORDS = 24
SENTINEL_WORD = 23
SENTINEL_BITS = 0xFFFFFFFF
META_CAT = 0x04000000
META_SLOT_MASK = 0x00FFFFFF
META_FEAT_SHIFT = 27
NODE_STRIDE = 256
GROUP_TREES = 32
MAX_TREES = 1024
HEADER_BYTES = 512
AGG_RF_MEAN = 0
AGG_GBDT_LOGISTIC = 1
AGG_IFOREST = 2
BLOB_VERSION = 2

_HEADER_FMT = "<8s" + "I" * 10 + "dd" + "Q" * 4 + "24f" + "24i" + "d"  # 296 bytes, padded to 512
_GROUP_FMT = "<8I"


@dataclass
class FlatForest:
    """A forest blob plus the host-side vocabulary needed to encode rows for it."""

    blob: bytes
    cat_features: list
    num_features: list
    categories: list  # per categorical feature: list[str], sorted as OneHotEncoder.categories_
    classes: list  # class labels in sklearn order, e.g. [0, 1]
    agg_mode: int
    n_trees: int
    max_depth: int
    total_nodes: int
    missing_codes: list = field(default_factory=list)  # per cat feature: code NaN maps to (the imputer's "missing") or -1
    none_codes: list = field(default_factory=list)  # per cat feature: code None maps to (a None category seen at fit) or -1

    @property
    def all_features(self):
        return list(self.cat_features) + list(self.num_features)

    def save(self, path: str) -> None:
        meta = dict(
            cat_features=self.cat_features,
            num_features=self.num_features,
            categories=self.categories,
            classes=self.classes,
            agg_mode=self.agg_mode,
            n_trees=self.n_trees,
            max_depth=self.max_depth,
            total_nodes=self.total_nodes,
            missing_codes=self.missing_codes,
            none_codes=self.none_codes,
        )
        np.savez_compressed(path, blob=np.frombuffer(self.blob, dtype=np.uint8), meta=np.array(json.dumps(meta)))

    @staticmethod
    def load(path: str) -> "FlatForest":
        with np.load(path) as z:
            meta = json.loads(str(z["meta"]))
            return FlatForest(blob=z["blob"].tobytes(), **meta)


def floor_to_f32(t64: np.ndarray) -> np.ndarray:
    """Largest float32 <= t64 (elementwise); +-inf pass through."""
    t64 = np.asarray(t64, dtype=np.float64)
    with np.errstate(over="ignore"):
        t32 = t64.astype(np.float32)
    too_big = t32.astype(np.float64) > t64
    t32[too_big] = np.nextafter(t32[too_big], np.float32(-np.inf))
    return t32


def _bfs_slots(left: np.ndarray, right: np.ndarray):
    """Breadth-first renumbering with adjacent siblings.  Returns (slot_of_node, depth)."""
    n = left.shape[0]
    slot = np.full(n, -1, dtype=np.int64)
    slot[0] = 0
    nxt = 1
    level = np.array([0], dtype=np.int64)
    depth = 0
    while True:
        internal = level[left[level] != -1]
        if internal.size == 0:
            break
        first = nxt + 2 * np.arange(internal.size, dtype=np.int64)
        slot[left[internal]] = first
        slot[right[internal]] = first + 1
        nxt += 2 * internal.size
        level = np.empty(2 * internal.size, dtype=np.int64)
        level[0::2] = left[internal]
        level[1::2] = right[internal]
        depth += 1
    assert nxt == n and (slot >= 0).all(), "tree has unreachable nodes"
    return slot, depth


def strict_upper_f32(t64: np.ndarray) -> np.ndarray:
    """t' = nextup(floor32(t64)): the float32 with  x <= t64  <=>  x < t'  for every finite float32 x."""
    return np.nextafter(floor_to_f32(t64), np.float32(np.inf))


def _flatten_tree(tree, col_word, col_cat_code, col_is_cat, leaf_value):
    """One sklearn ``Tree`` -> (T uint32[n], M uint32[n], LV float64[n_leaves], depth), slot-indexed."""
    left = tree.children_left.astype(np.int64)
    right = tree.children_right.astype(np.int64)
    n = left.shape[0]
    if n >= (1 << 24):
        raise NotImplementedError("tree too large for 24-bit slot offsets")
    slot, depth = _bfs_slots(left, right)
    T = np.zeros(n, dtype=np.uint32)
    M = np.zeros(n, dtype=np.uint32)
    is_leaf = left == -1
    internal = np.nonzero(~is_leaf)[0]
    leaves = np.nonzero(is_leaf)[0]

    # leaves: numbered in slot order; T = row of the leaf's payload in LV, M = self-loop
    leaf_order = leaves[np.argsort(slot[leaves])]
    leaf_id = np.arange(leaf_order.size, dtype=np.uint32)
    ls = slot[leaf_order]
    T[ls] = leaf_id
    M[ls] = ls.astype(np.uint32) | np.uint32(META_CAT) | np.uint32(SENTINEL_WORD << META_FEAT_SHIFT)
    LV = leaf_value[leaf_order].astype(np.float64)

    if internal.size:
        col = tree.feature[internal].astype(np.int64)
        thr = tree.threshold[internal].astype(np.float64)
        s = slot[internal]
        first = slot[left[internal]].astype(np.uint32)
        word = col_word[col].astype(np.uint32) << np.uint32(META_FEAT_SHIFT)
        cat = col_is_cat[col]
        t_words = strict_upper_f32(thr).view(np.uint32).copy()
        m_words = first | word
        if cat.any():
            # one-hot column x in {0, 1}:  x <= thr ?  x=0 -> (0 <= thr), x=1 -> (1 <= thr)
            zero_left = 0.0 <= thr
            one_left = 1.0 <= thr
            normal = cat & zero_left & ~one_left  # the only case sklearn produces (thr = 0.5)
            always_left = cat & zero_left & one_left
            always_right = cat & ~zero_left
            t_words[normal] = col_cat_code[col[normal]].astype(np.uint32)
            m_words[normal] |= np.uint32(META_CAT)
            t_words[always_left] = np.uint32(0x7FFFFFFF)  # never equals a category code
            m_words[always_left] |= np.uint32(META_CAT)
            # always second child: numeric test on the sentinel word (NaN bits): geu(NaN, t) is true
            t_words[always_right] = np.uint32(0)
            m_words[always_right] = first[always_right] | np.uint32(SENTINEL_WORD << META_FEAT_SHIFT)
        T[s] = t_words
        M[s] = m_words
    return T, M, LV, depth


def _assemble_blob(flat, agg, init_raw, denom, n_cat, n_num, impute, vocab, threshold=0.0):
    """Flattened trees ``[(T, M, LV, depth), ...]`` -> (blob bytes, max depth): groups of 32 interleaved trees."""
    n_trees = len(flat)
    if not (1 <= n_trees <= MAX_TREES):
        raise NotImplementedError(f"n_trees={n_trees} outside [1, {MAX_TREES}]")
    n_groups = (n_trees + GROUP_TREES - 1) // GROUP_TREES
    groups, chunks, off = [], [], 0
    for g in range(n_groups):
        members = flat[g * GROUP_TREES : (g + 1) * GROUP_TREES]
        n_slots = max(len(m[0]) for m in members)
        n_leaf = max(len(m[2]) for m in members)
        depth = max(m[3] for m in members)
        N = np.empty((n_slots, GROUP_TREES, 2), dtype=np.uint32)  # [slot][tree] -> (T, M)
        LV = np.zeros((n_leaf, GROUP_TREES), dtype=np.float64)
        # unused slots / stub trees: self-looping leaf with payload row 0 (value 0.0 for stubs)
        N[:, :, 0] = 0
        N[:, :, 1] = np.arange(n_slots, dtype=np.uint32)[:, None] | np.uint32(META_CAT | (SENTINEL_WORD << META_FEAT_SHIFT))
        for lane, (t, m, lv, _) in enumerate(members):
            N[: len(t), lane, 0] = t
            N[: len(m), lane, 1] = m
            LV[: len(lv), lane] = lv
        chunk = N.tobytes() + LV.tobytes()
        assert len(chunk) == (n_slots + n_leaf) * 256
        groups.append((off, len(chunk), n_slots, n_leaf, depth, len(members), 0, 0))
        chunks.append(chunk)
        off += len(chunk)

    max_depth = max(m[3] for m in flat)
    groups_off = HEADER_BYTES
    chunks_off = (groups_off + 32 * n_groups + 255) // 256 * 256
    total = chunks_off + off
    header = struct.pack(
        _HEADER_FMT,
        b"B2FOREST",
        BLOB_VERSION,
        HEADER_BYTES,
        agg,
        n_trees,
        n_groups,
        ROW_WORDS,
        n_cat,
        n_num,
        max_depth,
        0,
        init_raw,
        denom,
        groups_off,
        chunks_off,
        off,
        total,
        *np.asarray(impute, dtype=np.float32).tolist(),
        *np.asarray(vocab, dtype=np.int32).tolist(),
        float(threshold),
    )
    header = header + b"\0" * (HEADER_BYTES - len(header))
    table = b"".join(struct.pack(_GROUP_FMT, *g) for g in groups)
    pad = b"\0" * (chunks_off - groups_off - len(table))
    blob = header + table + pad + b"".join(chunks)
    assert len(blob) == total
    return blob, int(max_depth)


def _average_path_length(n):
    """c(n): average path length of an unsuccessful BST search over n points (Liu et al. 2008, eq. 1), with
    sklearn's conventions c(<=1) = 0, c(2) = 1 (``sklearn/ensemble/_iforest.py`` ``_average_path_length``)."""
    n = np.asarray(n, dtype=np.float64)
    out = np.zeros(n.shape, dtype=np.float64)
    out[n == 2] = 1.0
    big = n > 2
    out[big] = 2.0 * (np.log(n[big] - 1.0) + np.euler_gamma) - 2.0 * (n[big] - 1.0) / n[big]
    return out


def flatten_isolation_forest(detector, n_cat: int, n_num: int, vocab=None, threshold: float | None = None) -> bytes:
    """Fitted outlier detector -> forest blob with ``agg_mode = AGG_IFOREST`` over the classifier's encoded rows.

    ``detector`` is a fitted ``sklearn.ensemble.IsolationForest`` or an object carrying one as
    ``.isolationforest`` plus ``.threshold`` (alibi-detect's ``IForest``; the reference builds
    ``IForest(threshold=0.95).fit(df[NUMERIC_FEATURES].values)``, ``02-register-model.ipynb:232-233``, and calls
    ``.predict(df[numeric_features].values)`` per request, ``:339,344``).  Feature k of the detector is numeric
    feature k of the request, i.e. row word ``n_cat + k``.  What the GPU reproduces:

    * ``score = -decision_function(X) = 2 ** (-sum_t h_t(x) / (n_trees * c(max_samples))) + offset_`` with
      ``h_t(x) = depth(leaf) + c(n_node_samples[leaf])`` (``_iforest.py`` ``_compute_score_samples``);
    * ``is_outlier = score > threshold``.

    Splits compare float32 inputs with float64 thresholds exactly as the classifier's trees do.  NaN inputs are
    outside the contract: the reference's pinned scikit-learn 1.1.1 (``app/requirements.txt:14``) rejects them in
    ``IsolationForest.decision_function`` (ValueError -> HTTP 500) and ``B200Model`` does the same; newer sklearn
    routes them by a per-node random ``missing_go_to_left`` flag that the 8-byte node does not carry (on the GPU a
    NaN takes the second child at every split: the imputation table of this blob holds NaN).
    """
    iso = getattr(detector, "isolationforest", detector)
    if type(iso).__name__ != "IsolationForest":
        raise NotImplementedError(f"unsupported outlier detector {type(iso).__name__}")
    if threshold is None:
        threshold = getattr(detector, "threshold", None)
    if threshold is None:
        raise ValueError("an outlier threshold is required (alibi-detect IForest(threshold=...))")
    if iso.n_features_in_ != n_num:
        raise NotImplementedError(f"detector was fitted on {iso.n_features_in_} features, the request schema has {n_num} numerics")
    if n_cat + n_num > SENTINEL_WORD:
        raise NotImplementedError(f"at most {SENTINEL_WORD} raw features supported")
    flat = []
    for est, feats in zip(iso.estimators_, iso.estimators_features_):
        tree = est.tree_
        feats = np.asarray(feats, dtype=np.int64)
        col_word = n_cat + feats  # the tree's local column j is detector feature feats[j]
        none = np.zeros(col_word.shape[0], dtype=bool)
        # h(leaf) = number of edges from the root + c(training points that ended in the leaf)
        left, right = tree.children_left, tree.children_right
        depth = np.zeros(tree.node_count, dtype=np.float64)
        for i in range(tree.node_count):  # sklearn stores parents before children
            if left[i] != -1:
                depth[left[i]] = depth[right[i]] = depth[i] + 1.0
        payload = depth + _average_path_length(tree.n_node_samples)
        flat.append(_flatten_tree(tree, col_word, np.zeros_like(col_word), none, payload))
    denom = float(len(flat)) * float(_average_path_length(np.array([iso._max_samples]))[0])
    if not denom > 0.0:
        raise NotImplementedError("isolation forest fitted on a single sample")
    impute = np.zeros(ROW_WORDS, dtype=np.float32)
    impute[n_cat : n_cat + n_num] = np.nan
    v = np.zeros(ROW_WORDS, dtype=np.int32)
    if vocab is not None:
        v[:n_cat] = np.asarray(vocab, dtype=np.int32)[:n_cat]
    blob, _ = _assemble_blob(flat, AGG_IFOREST, float(iso.offset_), denom, n_cat, n_num, impute, v, threshold=float(threshold))
    return blob


def _describe_preprocessor(pre):
    """ColumnTransformer of the reference shape -> column maps for its output matrix."""
    cat_cols, num_cols, ohe, num_imputer, cat_imputer = None, None, None, None, None
    for name, trans, cols in pre.transformers_:
        if trans == "drop" or (isinstance(trans, str) and trans == "passthrough"):
            continue
        steps = dict(trans.named_steps) if hasattr(trans, "named_steps") else {"only": trans}
        kinds = {type(s).__name__ for s in steps.values()}
        if "OneHotEncoder" in kinds:
            cat_cols = list(cols)
            ohe = next(s for s in steps.values() if type(s).__name__ == "OneHotEncoder")
            cat_imputer = next((s for s in steps.values() if type(s).__name__ == "SimpleImputer"), None)
            cat_name = name
        else:
            num_cols = list(cols)
            num_imputer = next((s for s in steps.values() if type(s).__name__ == "SimpleImputer"), None)
            num_name = name
    if ohe is None or num_cols is None or num_imputer is None:
        raise NotImplementedError("expected ColumnTransformer([categorical: imputer+OneHotEncoder, numeric: imputer])")
    if ohe.handle_unknown != "ignore" or getattr(ohe, "drop", None) is not None:
        raise NotImplementedError("OneHotEncoder must use handle_unknown='ignore' and drop=None")
    if len(cat_cols) + len(num_cols) > SENTINEL_WORD:
        raise NotImplementedError(f"at most {SENTINEL_WORD} raw features supported")
    # a training column that held None / NaN gives OneHotEncoder a non-string category (sorted last); keep its
    # position under a placeholder no request string can equal, and remember it for the encoder
    NONE_TOKEN = "\x00<none>"
    categories = [[c if isinstance(c, str) else NONE_TOKEN for c in cats] for cats in ohe.categories_]
    n_cat, n_num = len(cat_cols), len(num_cols)
    cat_slice = pre.output_indices_[cat_name]
    num_slice = pre.output_indices_[num_name]
    n_out = max(cat_slice.stop, num_slice.stop)
    col_word = np.zeros(n_out, dtype=np.int64)
    col_code = np.zeros(n_out, dtype=np.int64)
    col_is_cat = np.zeros(n_out, dtype=bool)
    j = cat_slice.start
    for f, cats in enumerate(categories):
        for c in range(len(cats)):
            col_word[j], col_code[j], col_is_cat[j] = f, c, True
            j += 1
    assert j == cat_slice.stop
    for k in range(n_num):
        col_word[num_slice.start + k] = n_cat + k
    medians = np.asarray(num_imputer.statistics_, dtype=np.float64)
    fill = getattr(cat_imputer, "fill_value", None) if cat_imputer is not None else None
    # NaN is imputed to the constant `fill` ("missing") before encoding; None is left alone by SimpleImputer and
    # only matches a None category seen at fit time
    missing_codes = [cats.index(fill) if (fill is not None and fill in cats) else -1 for cats in categories]
    none_codes = [cats.index(NONE_TOKEN) if NONE_TOKEN in cats else -1 for cats in categories]
    return cat_cols, num_cols, categories, medians, col_word, col_code, col_is_cat, missing_codes, none_codes


def flatten_pipeline(pipeline) -> FlatForest:
    """Fitted reference-style Pipeline -> FlatForest."""
    pre = pipeline.named_steps["preprocessor"]
    clf = pipeline.named_steps["classifier"]
    cat_cols, num_cols, categories, medians, col_word, col_code, col_is_cat, missing_codes, none_codes = _describe_preprocessor(pre)
    n_cat, n_num = len(cat_cols), len(num_cols)

    kind = type(clf).__name__
    if kind == "RandomForestClassifier":
        if len(clf.classes_) != 2 or clf.n_outputs_ != 1:
            raise NotImplementedError("binary single-output RandomForestClassifier only")
        agg = AGG_RF_MEAN
        trees = [e.tree_ for e in clf.estimators_]
        init_raw, denom = 0.0, float(len(trees))

        def leaf_values(t):
            v = t.value[:, 0, :]
            s = v.sum(axis=1)
            s = np.where(s == 0.0, 1.0, s)
            return v[:, 1] / s  # class-1 fraction (normalised for sklearn < 1.4 weighted counts)

    elif kind == "GradientBoostingClassifier":
        if len(clf.classes_) != 2:
            raise NotImplementedError("binary GradientBoostingClassifier only")
        agg = AGG_GBDT_LOGISTIC
        trees = [e.tree_ for e in clf.estimators_[:, 0]]
        if isinstance(clf.init_, str) and clf.init_ == "zero":
            init_raw = 0.0
        elif type(clf.init_).__name__ == "DummyClassifier":
            eps = np.finfo(np.float64).eps
            p1 = float(np.clip(clf.init_.class_prior_[1], eps, 1 - eps))
            init_raw = float(np.log(p1 / (1 - p1)))
        else:
            raise NotImplementedError("GBDT init estimator must be the default prior or 'zero'")
        denom = 1.0
        lr = np.float64(clf.learning_rate)

        def leaf_values(t):
            return lr * t.value[:, 0, 0].astype(np.float64)

    else:
        raise NotImplementedError(f"unsupported classifier {kind}")

    flat = [_flatten_tree(t, col_word, col_code, col_is_cat, leaf_values(t)) for t in trees]
    impute = np.zeros(ROW_WORDS, dtype=np.float32)
    with np.errstate(over="ignore"):
        impute[n_cat : n_cat + n_num] = medians.astype(np.float32)
    vocab = np.zeros(ROW_WORDS, dtype=np.int32)
    vocab[:n_cat] = [len(c) for c in categories]
    blob, max_depth = _assemble_blob(flat, agg, init_raw, denom, n_cat, n_num, impute, vocab)
    n_trees = len(trees)
    return FlatForest(
        blob=blob,
        cat_features=[str(c) for c in cat_cols],
        num_features=[str(c) for c in num_cols],
        categories=categories,
        classes=[c.item() if hasattr(c, "item") else c for c in clf.classes_],
        agg_mode=agg,
        n_trees=n_trees,
        max_depth=int(max_depth),
        total_nodes=int(sum(len(m[0]) for m in flat)),
        missing_codes=missing_codes,
        none_codes=none_codes,
    )


def parse_header(blob: bytes) -> dict:
    """Decode the fixed header + group table (host-side mirror of ``forest_blob.h``)."""
    f = struct.unpack_from(_HEADER_FMT, blob, 0)
    keys = ["magic", "version", "header_bytes", "agg_mode", "n_trees", "n_groups", "row_words", "n_cat", "n_num",
            "max_depth", "reserved0", "init_raw", "denom", "groups_off", "chunks_off", "chunks_bytes", "total_bytes"]
    h = dict(zip(keys, f[:17]))
    h["impute"] = np.array(f[17:41], dtype=np.float32)
    h["vocab"] = np.array(f[41:65], dtype=np.int32)
    h["threshold"] = f[65]
    gk = ["chunk_off", "chunk_bytes", "n_slots", "n_leaf_slots", "depth", "n_trees"]
    h["groups"] = [
        dict(zip(gk, struct.unpack_from(_GROUP_FMT, blob, h["groups_off"] + 32 * g)[:6])) for g in range(h["n_groups"])
    ]
    return h
