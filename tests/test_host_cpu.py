"""Host-side logic on a CPU-only box: flattener + blob format (through a numpy emulation of the kernel),
row encoder, C-ABI surface.  No compute call touches the GPU here."""

import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- C ABI surface
def _declared_functions():
    src = open(os.path.join(ROOT, "include", "b2f.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2f_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from databricks_kubernetes_mlops_poc_b200 import _cabi

    lib = _cabi.load_library()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b2f.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes prototype"
    assert set(_cabi.SIGNATURES) == set(names)
    assert lib.b2f_version().startswith(b"b200forest")


def test_no_device_means_no_result(rf100d6):
    """No CPU fallback: on a box without a GPU model creation fails loudly (on a GPU box this is skipped)."""
    from conftest import has_gpu
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200._cabi import B2FError
    from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine

    if has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(B2FError, match="no CUDA device|failed"):
        ForestEngine(flatten.flatten_pipeline(rf100d6), 0)


def test_drift_detector_needs_the_device(curated, tmp_path):
    """The drift detector has no CPU path either: reference statistics can be prepared and saved anywhere, scoring
    (and opening a saved detector) needs the GPU."""
    from conftest import has_gpu
    from oracle import reference_pipeline as rp

    from databricks_kubernetes_mlops_poc_b200._cabi import B2FError
    from databricks_kubernetes_mlops_poc_b200.drift import TabularDrift

    ref = curated[rp.FEATURES].iloc[:2000]
    det = TabularDrift(ref, rp.CATEGORICAL_FEATURES, device=None)
    assert det.features == rp.FEATURES and len(det.ref_sorted) == 14 and len(det.ref_cats) == 9
    assert (np.diff(det.ref_sorted[rp.NUMERIC_FEATURES[0]]) >= 0).all()
    assert int(det.ref_counts["sex"].sum()) == 2000
    # host half of a request: category indices against the reference categories, counts of values outside them
    big = curated[rp.FEATURES].iloc[2000:2600].copy().reset_index(drop=True)
    big.loc[[5, 9, 11], "sex"] = ["zz_unseen", "aa_unseen", "zz_unseen"]
    for batch in (big, big.iloc[:40]):  # vectorised path (n > 128) and request-sized path agree with the definition
        x, codes, new_off, newc = det.encode_batch(batch)
        assert x.shape == (14, len(batch)) and codes.shape == (9, len(batch))
        for c, name in enumerate(det.cat_features):
            vals = batch[name].astype(str).to_numpy()
            cats = det.ref_cats[name].tolist()
            want = np.array([cats.index(v) if v in cats else -1 for v in vals])
            assert (codes[c] == want).all()
            outside = sorted(set(vals[want < 0].tolist()))
            assert newc[new_off[c]:new_off[c + 1]].tolist() == [int((vals == v).sum()) for v in outside]
        assert (x[0] == batch[det.num_features[0]].to_numpy()).all()
        assert new_off[-1] == len(newc) >= 2 and {1, 2} <= set(newc.tolist())
    # missing category values (outside the reference's contract) count as one category "nan" on both paths
    holes = big.copy()
    holes["education"] = holes["education"].astype(object)
    holes.loc[[1, 2], "education"] = [None, np.nan]
    c = det.cat_features.index("education")
    for batch in (holes, holes.iloc[:20]):
        _, codes, new_off, newc = det.encode_batch(batch)
        assert (codes[c][[1, 2]] == -1).all() and 2 in newc[new_off[c]:new_off[c + 1]].tolist()
        assert (codes[c][[0, 3]] == det.ref_cats["education"].tolist().index(batch["education"].iloc[0])).all()
    det.save(str(tmp_path / "d.npz"))
    with pytest.raises(B2FError, match="no CPU fallback"):
        det.statistics(ref.iloc[:5])
    if has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(B2FError, match="b2f_drift_create"):
        TabularDrift(ref, rp.CATEGORICAL_FEATURES, device=0)
    with pytest.raises(B2FError, match="b2f_drift_create"):
        TabularDrift.load(str(tmp_path / "d.npz"), device=0)


def test_mlflow_shim_resolves_to_the_b200_loader(monkeypatch):
    """with the shim directory on the path, the reference's `mlflow.pyfunc.load_model(dir)` call
    (app/main.py:26-28) lands in this package's loader; the shim is opt-in by path and not imported otherwise."""
    import importlib
    import sys

    import databricks_kubernetes_mlops_poc_b200 as pkg

    shim = os.path.join(os.path.dirname(pkg.__file__), "shim")
    monkeypatch.syspath_prepend(shim)
    for name in [m for m in sys.modules if m == "mlflow" or m.startswith("mlflow.")]:
        monkeypatch.delitem(sys.modules, name)
    mlflow = importlib.import_module("mlflow")
    assert mlflow.__file__.startswith(shim) and callable(mlflow.pyfunc.load_model)
    seen = {}
    monkeypatch.setattr(pkg, "load_model", lambda path: seen.setdefault("path", path) or "model")
    mlflow.pyfunc.load_model("/models/credit_default")
    assert seen["path"] == "/models/credit_default"
    for name in [m for m in sys.modules if m == "mlflow" or m.startswith("mlflow.")]:
        monkeypatch.delitem(sys.modules, name)


def test_missing_library_fails_loudly(tmp_path):
    from databricks_kubernetes_mlops_poc_b200 import _cabi

    with pytest.raises(_cabi.B2FError, match="no CPU fallback"):
        _cabi.load_library(str(tmp_path / "nope.so"))


def test_moments_merge_is_chan(rf100d6):
    from databricks_kubernetes_mlops_poc_b200.engine import moments_merge

    rng = np.random.default_rng(0)
    x = rng.normal(5e4, 7e3, size=(5000, 24))
    cuts = [0, 700, 701, 3000, 5000]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        seg = x[a:b]
        parts.append(np.stack([np.full(24, b - a, float), seg.mean(0), ((seg - seg.mean(0)) ** 2).sum(0)], axis=1))
    parts.append(np.zeros((24, 3)))  # an empty shard
    got = moments_merge(np.stack(parts))
    assert np.allclose(got[:, 0], 5000)
    assert np.allclose(got[:, 1], x.mean(0), rtol=1e-13)
    assert np.allclose(got[:, 2] / 5000, x.var(0), rtol=1e-11)


# ----------------------------------------------------------------------------- flattener / blob
def _emulate(pipe, df):
    from blob_walk import walk_blob
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.engine import validate_blob

    flat = flatten.flatten_pipeline(pipe)
    validate_blob(flat.blob)
    return walk_blob(flat.blob, RowEncoder(flat).encode_frame(df)), flat


def test_blob_semantics_match_oracle(curated, inference, adversarial, rf100d6, gbdt_small):
    from oracle import reference_pipeline as rp

    for pipe in (rf100d6, gbdt_small):
        for df in (curated.iloc[:1500], inference, adversarial):
            (p, l), _ = _emulate(pipe, df)
            want_p, want_l = rp.oracle_predict(pipe, df)
            assert np.abs(p - want_p).max() < 1e-14 and (l == want_l).all()


def _iforest_rows(rf100d6, df):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder

    return RowEncoder(flatten.flatten_pipeline(rf100d6)).encode_frame(df)


def test_isolation_forest_blob_matches_sklearn(curated, inference, iforest, iforest_edges, rf100d6):
    """the outlier detector as a second forest blob over the classifier's encoded rows.
    score = -decision_function (alibi-detect IForest.score), flag = score > threshold."""
    from blob_walk import walk_blob
    from oracle import reference_pipeline as rp

    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.engine import validate_blob

    for thr in (0.95, 0.0, 0.04):
        blob = flatten.flatten_isolation_forest(iforest, 9, 14, threshold=thr)
        validate_blob(blob)
        h = flatten.parse_header(blob)
        assert h["agg_mode"] == flatten.AGG_IFOREST and h["n_trees"] == 100 and h["threshold"] == thr and h["init_raw"] == iforest.offset_
        for df in (curated.iloc[:3000], inference, iforest_edges):
            score, flag = walk_blob(blob, _iforest_rows(rf100d6, df))
            want = -iforest.decision_function(df[rp.NUMERIC_FEATURES].to_numpy())
            assert np.abs(score - want).max() < 1e-14
            assert (flag == (want > thr)).all()
            if thr == 0.95:
                assert flag.sum() == 0  # the reference's threshold can never fire (score <= 0.5)
            if thr == 0.0 and len(df) > 1000:
                assert 0 < flag.sum() < len(df)


@pytest.mark.parametrize("params", [dict(n_estimators=33, max_samples=64, random_state=1), dict(n_estimators=7, max_features=5, random_state=2),
                                    dict(n_estimators=40, max_samples=0.5, bootstrap=True, random_state=3),
                                    dict(n_estimators=1, max_samples=2, random_state=4), dict(n_estimators=65, max_samples=3, random_state=5),
                                    dict(n_estimators=4, max_samples=4096, contamination=0.1, random_state=6)])
def test_isolation_forest_variants(curated, rf100d6, params):
    """Feature sub-sampling (estimators_features_), bootstrap and other tree sizes; alibi-style wrapper object."""
    from types import SimpleNamespace

    from blob_walk import walk_blob
    from sklearn.ensemble import IsolationForest

    from oracle import reference_pipeline as rp

    from databricks_kubernetes_mlops_poc_b200 import flatten

    X = curated[rp.NUMERIC_FEATURES].to_numpy()
    iso = IsolationForest(**params).fit(X[:5000])
    blob = flatten.flatten_isolation_forest(SimpleNamespace(isolationforest=iso, threshold=0.01), 9, 14)
    df = curated.iloc[5000:7000]
    score, flag = walk_blob(blob, _iforest_rows(rf100d6, df))
    want = -iso.decision_function(df[rp.NUMERIC_FEATURES].to_numpy())
    assert np.abs(score - want).max() < 1e-14 and (flag == (want > 0.01)).all()
    with pytest.raises(NotImplementedError):
        flatten.flatten_isolation_forest(iso, 9, 13, threshold=0.5)  # schema mismatch
    with pytest.raises(ValueError):
        flatten.flatten_isolation_forest(iso, 9, 14)  # no threshold anywhere


def test_blob_layout_invariants(rf100d6):
    from databricks_kubernetes_mlops_poc_b200 import flatten

    flat = flatten.flatten_pipeline(rf100d6)
    h = flatten.parse_header(flat.blob)
    assert h["magic"] == b"B2FOREST" and h["version"] == flatten.BLOB_VERSION
    assert h["n_trees"] == 100 and h["n_groups"] == 4 and h["n_cat"] == 9 and h["n_num"] == 14
    assert h["denom"] == 100.0 and h["total_bytes"] == len(flat.blob)
    assert h["chunks_off"] % 256 == 0
    assert [g["n_trees"] for g in h["groups"]] == [32, 32, 32, 4]
    off = 0
    for g in h["groups"]:
        assert g["chunk_off"] == off and g["chunk_bytes"] == (g["n_slots"] + g["n_leaf_slots"]) * 256
        off += g["chunk_bytes"]
    assert off == h["chunks_bytes"] <= 227 * 1024  # the pinned 100 x depth-6 forest is shared-memory resident
    assert (h["vocab"][:9] == [2, 7, 4, 10, 10, 10, 10, 9, 9]).all()
    rt = flatten.FlatForest.load  # save / load round trip
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        flat.save(os.path.join(d, "f.npz"))
        assert rt(os.path.join(d, "f.npz")).blob == flat.blob


def test_threshold_conversion_is_exact():
    from databricks_kubernetes_mlops_poc_b200.flatten import floor_to_f32, strict_upper_f32

    rng = np.random.default_rng(3)
    t = np.concatenate([rng.normal(0, 1e4, 2000), rng.normal(0, 1e-3, 500), [0.0, -0.0, 0.5, 0.27500000596046448, 1e39, -1e39, 3.4028234e38]])
    x = np.concatenate([rng.normal(0, 1e4, 3000).astype(np.float32), t.astype(np.float32, casting="unsafe")[np.isfinite(t.astype(np.float32))],
                        np.float32([0.0, -0.0, 1e-45, -1e-45, 3.4e38, -3.4e38])])
    f = floor_to_f32(t)
    u = strict_upper_f32(t)
    want = x[:, None].astype(np.float64) <= t[None, :]  # sklearn: float32 x vs float64 threshold
    assert (want == (x[:, None] <= f[None, :])).all()
    assert (want == (x[:, None] < u[None, :])).all()


def test_malformed_blobs_are_rejected(rf100d6):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200._cabi import B2FError
    from databricks_kubernetes_mlops_poc_b200.engine import validate_blob

    blob = flatten.flatten_pipeline(rf100d6).blob
    h = flatten.parse_header(blob)
    validate_blob(blob)
    for mutate in (
        lambda b: b.__setitem__(0, 0),  # magic
        lambda b: b.__setitem__(8, 9),  # version
        lambda b: b.__setitem__(slice(h["chunks_off"] + 4, h["chunks_off"] + 8), (0x00FFFFF0).to_bytes(4, "little")),  # child out of range
        lambda b: b.__setitem__(slice(h["chunks_off"] + 4, h["chunks_off"] + 8), (31 << 27 | 1).to_bytes(4, "little")),  # row word 31
    ):
        bad = bytearray(blob)
        mutate(bad)
        with pytest.raises(B2FError):
            validate_blob(bytes(bad))
    with pytest.raises(B2FError):
        validate_blob(blob[:-1])
    with pytest.raises(B2FError):
        validate_blob(b"short")


def test_unsupported_models_are_refused(curated):
    from sklearn.ensemble import RandomForestClassifier

    from databricks_kubernetes_mlops_poc_b200 import flatten, training

    pipe = training.make_pipeline("rf", n_estimators=3, max_depth=2, random_state=0)
    y3 = (np.arange(500) % 3)
    pipe.fit(curated[flatten_features()].iloc[:500], y3)  # 3 classes
    with pytest.raises(NotImplementedError):
        flatten.flatten_pipeline(pipe)
    assert isinstance(pipe.named_steps["classifier"], RandomForestClassifier)


def flatten_features():
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES

    return ALL_FEATURES


# ----------------------------------------------------------------------------- encoder
def test_encoder_matches_reference_lookup(curated, inference, adversarial, rf100d6):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from oracle import treewalk as tw

    flat = flatten.flatten_pipeline(rf100d6)
    enc = RowEncoder(flat)
    dump = tw.dump_pipeline(rf100d6)
    for df in (curated.iloc[:800], inference, adversarial):
        rows = enc.encode_frame(df)
        codes, nums = tw.encode_frame(dump, df)
        assert rows.shape == (len(df), 24) and rows.dtype == np.uint32
        assert (rows.view(np.int32)[:, :9] == codes).all()
        got = rows.view(np.float32)[:, 9:23]
        want = nums.astype(np.float32)
        assert ((got == want) | (np.isnan(got) & np.isnan(want))).all()
        assert (rows[:, 23] == 0).all()
        shuffled = df[list(reversed([c for c in df.columns]))]  # column order must not matter
        assert (enc.encode_frame(shuffled) == rows).all()


def test_encoder_errors(curated, rf100d6):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES

    enc = RowEncoder(flatten.flatten_pipeline(rf100d6))
    df = curated[ALL_FEATURES].iloc[:3].copy()
    for bad in (np.inf, -np.inf, 1e39):
        d = df.copy()
        d.loc[d.index[1], "bill_amount_3"] = bad
        with pytest.raises(ValueError, match="infinity or a value too large"):
            enc.encode_frame(d)
    with pytest.raises(KeyError):
        enc.encode_frame(df.drop(columns=["sex"]))
    d = df.copy()
    d["age"] = d["age"].astype(object)
    d.loc[d.index[0], "age"] = "forty"
    with pytest.raises((ValueError, TypeError)):
        enc.encode_frame(d)
    assert enc.encode_frame(df.iloc[:0]).shape == (0, 24)


def test_schema_is_wire_compatible():
    """Field names, order, types and defaults of the reference's pydantic models (app/model.py:8-70)."""
    from databricks_kubernetes_mlops_poc_b200 import schema

    fields = schema.LoanApplicant.model_fields
    assert list(fields) == schema.ALL_FEATURES and len(fields) == 23
    assert [fields[n].annotation for n in schema.CATEGORICAL_FEATURES] == [str] * 9
    assert [fields[n].annotation for n in schema.NUMERIC_FEATURES] == [float] * 14
    row = schema.LoanApplicant()
    assert row.sex == "male" and row.repayment_status_5 == "no_delay" and row.age == 18000.0 and row.payment_amount_6 == 805.65
    assert schema.LoanApplicant.model_validate({}).model_dump() == schema.sample_request()[0]
    with pytest.raises(Exception):
        schema.LoanApplicant.model_validate({"sex": 3})
    out = schema.ModelOutput.model_validate(
        {"predictions": [0.5], "outliers": [0], "feature_drift_batch": {n: 0.0 for n in schema.ALL_FEATURES}})
    assert out.model_dump()["outliers"] == [0.0]
    with pytest.raises(Exception):
        schema.ModelOutput.model_validate({"predictions": [0.5], "outliers": [0], "feature_drift_batch": {"sex": 0.0}})


def test_packed_rows_are_lossless(curated, adversarial, rf100d6):
    """B2F_ROWS_PACKED64: nine 7-bit (code + 1) fields + 14 float32 -> unpacking gives back the 96-byte row."""
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder

    enc = RowEncoder(flatten.flatten_pipeline(rf100d6))
    assert enc.packed_ok
    for df in (curated.iloc[:2000], adversarial):
        rows = enc.encode_frame(df)
        pk = enc.encode_frame_packed(df)
        assert pk.shape == (len(df), 16) and pk.dtype == np.uint32
        word = pk[:, 0].astype(np.uint64) | (pk[:, 1].astype(np.uint64) << np.uint64(32))
        for j in range(9):
            code = ((word >> np.uint64(7 * j)) & np.uint64(0x7F)).astype(np.int64) - 1
            assert (code == rows.view(np.int32)[:, j]).all()
        assert (pk[:, 2:16] == rows[:, 9:23]).all()  # numerics bit-identical (NaN payloads included)
        assert (word >> np.uint64(63) == 0).all()


@pytest.mark.parametrize("seed", range(6))
def test_random_forest_shapes_through_the_blob(curated, adversarial, seed):
    """Property check over model shapes: random (n_trees, depth, criterion, training subset, GBDT / RF) ->
    flatten -> blob validates -> kernel-semantics emulation == the library, labels exact."""
    from oracle import reference_pipeline as rp

    rng = np.random.default_rng(100 + seed)
    n_trees = int(rng.choice([1, 2, 31, 32, 33, 65, 130]))
    depth = int(rng.choice([1, 2, 3, 5, 9, 14]))
    lo = int(rng.integers(0, 20000))
    train = curated.iloc[lo : lo + int(rng.choice([200, 1500, 4000]))]
    if seed % 3 == 2:
        pipe = rp.fit_gbdt_pipeline(train, train[rp.TARGET].to_numpy(), dict(n_estimators=min(n_trees, 40), max_depth=min(depth, 5), random_state=seed))
    else:
        pipe = rp.make_classifier_pipeline(dict(n_estimators=n_trees, max_depth=depth, criterion=["gini", "entropy"][seed % 2], random_state=seed))
        pipe.fit(train[rp.FEATURES], train[rp.TARGET].to_numpy())
    probe = curated.iloc[25000:25400]
    for df in (probe, adversarial.iloc[:200]):
        (p, l), flat = _emulate(pipe, df)
        want_p, want_l = rp.oracle_predict(pipe, df)
        assert np.abs(p - want_p).max() < 1e-13 and (l == want_l).all()
    assert flat.n_trees == len(pipe.named_steps["classifier"].estimators_)


@pytest.mark.parametrize("hole", [None, np.nan])
def test_training_vocabulary_with_missing_values(curated, hole):
    """Missing values at FIT time change the vocabulary: NaN is imputed to the constant "missing"
    (SimpleImputer, 01-train-model.ipynb:200) and becomes a real category; None is left alone by the imputer
    and becomes OneHotEncoder's None category.  Requests must then map NaN / None / "missing" / unknown
    exactly as the library does."""
    from oracle import reference_pipeline as rp

    train = curated.iloc[:3000].copy()
    col = train["education"].astype(object)
    col.iloc[::7] = hole
    train["education"] = col
    pipe = rp.make_classifier_pipeline(dict(n_estimators=20, max_depth=6, random_state=0))
    pipe.fit(train[rp.FEATURES], train[rp.TARGET].to_numpy())
    probe = curated.iloc[5000:5300].copy()
    c2 = probe["education"].astype(object)
    c2.iloc[::3] = None
    c2.iloc[1::5] = np.nan
    c2.iloc[2::9] = "missing"
    c2.iloc[3::11] = "unheard_of"
    probe["education"] = c2
    (p, l), flat = _emulate(pipe, probe)
    want_p, want_l = rp.oracle_predict(pipe, probe)
    if hole is None:
        assert flat.none_codes[1] >= 0 and flat.missing_codes[1] == -1
    else:
        assert "missing" in flat.categories[1] and flat.missing_codes[1] >= 0
    assert np.abs(p - want_p).max() < 1e-13 and (l == want_l).all()
    (p1, l1), _ = _emulate(pipe, probe.iloc[:40])  # small-request encoder path
    assert np.abs(p1 - want_p[:40]).max() < 1e-13 and (l1 == want_l[:40]).all()


def test_native_encoder_matches_portable_encoder(curated, inference, rf100d6):
    """csrc/row_encoder.h (Arrow string buffers + float64 columns -> rows, C++ threads) == encode.py's portable
    path, for both row layouts, on sliced / reordered frames, nulls, unknown strings, integer numerics."""
    import pandas as pd

    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES

    enc = RowEncoder(flatten.flatten_pipeline(rf100d6))
    base = curated[ALL_FEATURES].iloc[:9000].copy()
    rng = np.random.default_rng(4)
    for name in ("education", "repayment_status_3"):
        col = base[name].astype(object)
        col[rng.random(len(base)) < 0.05] = "never_seen"
        col[rng.random(len(base)) < 0.03] = None
        base[name] = col.astype("str") if False else pd.array(col, dtype="str")  # Arrow-backed string column with nulls
    base["age"] = base["age"].astype(np.int64)
    num = base["bill_amount_2"].to_numpy(copy=True)
    num[::17] = np.nan
    base["bill_amount_2"] = num
    frames = [base, base.iloc[1234:7777], base[list(reversed(ALL_FEATURES))], inference]
    for df in frames:
        assert len(df) > RowEncoder.SMALL_BATCH or df is inference
        want = np.empty((len(df), 24), dtype=np.uint32)
        enc2 = RowEncoder(flatten.flatten_pipeline(rf100d6))
        enc2._native_failed = True  # portable path only
        enc2.encode_frame(df, out=want)
        got = np.zeros_like(want)
        used = enc._encode_native(df, got, packed=False)
        assert used, "native path should accept string-dtype columns"
        assert (got == want).all()
        pk = np.zeros((len(df), 16), dtype=np.uint32)
        assert enc._encode_native(df, pk, packed=True)
        assert (pk == enc2.pack_rows(want)).all()
        assert (enc.encode_frame_packed(df) == pk).all() and (enc.encode_frame(df) == want).all()
    bad = base.copy()
    bad.loc[bad.index[5000], "payment_amount_1"] = np.inf
    with pytest.raises(ValueError, match="infinity or a value too large"):
        enc.encode_frame(bad)
    bad.loc[bad.index[5000], "payment_amount_1"] = 1e39
    with pytest.raises(ValueError, match="infinity or a value too large"):
        enc.encode_frame_packed(bad)
    obj = base.copy()
    obj["sex"] = obj["sex"].astype(object)  # object columns take the portable path (None vs NaN semantics)
    assert not enc._encode_native(obj, np.zeros((len(obj), 24), dtype=np.uint32), packed=False)
    assert (enc.encode_frame(obj) == enc.encode_frame(base)).all()
