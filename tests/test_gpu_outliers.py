"""GPU parity for the outlier detector: the reference's
``IForest(threshold=0.95)`` = sklearn IsolationForest (02-register-model.ipynb:232-233,339,344) as a second
forest blob walked by the same CUDA kernels over the same encoded rows.

Oracle: the installed sklearn itself, ``score = -IsolationForest.decision_function(X)`` and
``is_outlier = score > threshold`` (what alibi-detect 0.12.0's ``IForest.predict`` computes; the wrapper is not
installed here).  Bar: flags bit-exact, |d score| <= 1e-12 (float64 output) / 2e-7 (float32 output)."""

import os
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 2e-7


def _num(df):
    from oracle import reference_pipeline as rp

    return df[rp.NUMERIC_FEATURES].to_numpy()


def _with_kernel(kernel, fn):
    old = os.environ.get("B2F_KERNEL")
    try:
        if kernel:
            os.environ["B2F_KERNEL"] = kernel
        return fn()
    finally:
        if old is None:
            os.environ.pop("B2F_KERNEL", None)
        else:
            os.environ["B2F_KERNEL"] = old


@pytest.mark.parametrize("kernel", ["warp", "tile", "split"])
def test_isolation_forest_alone(curated, inference, iforest, iforest_edges, rf100d6, kernel):
    """An isolation-forest blob is a model of its own: every predict entry point returns (score, flag)."""
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine

    enc = RowEncoder(flatten.flatten_pipeline(rf100d6))
    thr = 0.0
    blob = flatten.flatten_isolation_forest(iforest, 9, 14, threshold=thr)
    eng = _with_kernel(kernel, lambda: ForestEngine(blob, 0))
    try:
        assert eng.info()["agg"] == "iforest" and eng.info()["n_trees"] == 100
        frames = [curated.iloc[:2000] if kernel == "split" else curated, inference, iforest_edges]
        for df in frames:
            want = -iforest.decision_function(_num(df))
            rows = enc.encode_frame(df)
            for r in (rows, enc.pack_rows(rows)):
                s64, f64 = eng.predict_rows(r, np.float64)
                s32, f32 = eng.predict_rows(r, np.float32)
                assert np.abs(s64 - want).max() <= TOL64
                assert np.abs(s32.astype(np.float64) - want).max() <= TOL32
                assert (f64 == (want > thr)).all() and (f32 == f64).all()
        info = eng.info()
        assert info["launches"] > 0
        if kernel == "tile":
            assert info["launches_tile"] == info["launches"]
        if kernel == "split":
            assert info["launches_split"] == info["launches"]
    finally:
        eng.close()


def test_predict_full_matches_both_oracles(curated, iforest, iforest_edges, rf100d6, rf500d8):
    """b2f_predict_full: classifier + outlier forest on one H2D copy of the rows, 24-byte records back; every
    batch-size regime (split / warp / tile kernels, single chunk and the pipelined chunk plan)."""
    from oracle import reference_pipeline as rp

    from databricks_kubernetes_mlops_poc_b200 import flatten, training
    from databricks_kubernetes_mlops_poc_b200._cabi import SCORED_FULL_DTYPE, B2FError
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine

    assert SCORED_FULL_DTYPE.itemsize == 24
    for pipe in (rf100d6, rf500d8):
        flat = flatten.flatten_pipeline(pipe)
        enc = RowEncoder(flat)
        eng = ForestEngine(flat, 0)
        try:
            rows = enc.encode_frame(curated)
            with pytest.raises(B2FError, match="no outlier forest"):
                eng.predict_full(rows[:10])
            with pytest.raises(B2FError, match="B2F_AGG_IFOREST"):
                eng.attach_outlier_forest(flat.blob)  # a classifier blob is not an outlier forest
            thr = 0.02
            eng.attach_outlier_forest(flatten.flatten_isolation_forest(iforest, 9, 14, threshold=thr))
            assert eng.info()["outlier_trees"] == 100
            want_p, want_l = rp.oracle_predict(pipe, curated)
            want_s = -iforest.decision_function(_num(curated))
            for n in (0, 1, 7, 300, 4096, 16384, 24577, 30000):
                for r in (rows[:n], enc.pack_rows(rows[:n])):
                    before = eng.info()["launches"]
                    out = eng.predict_full(r)
                    assert out.shape == (n,)
                    if n == 0:
                        continue
                    assert eng.info()["launches"] >= before + 2  # both forests were walked on the GPU
                    assert np.abs(out["proba1"] - want_p[:n]).max() <= TOL64 and (out["label"] == want_l[:n]).all()
                    assert np.abs(out["outlier_score"].astype(np.float64) - want_s[:n]).max() <= TOL32
                    assert (out["is_outlier"] == (want_s[:n] > thr)).all()
            # a batch large enough for the pipelined chunk plan and the tile kernel
            _, codes, nums = training.synth_arrays(curated, 70001, seed=5)
            big = enc.encode_arrays_packed(codes, nums)
            out = eng.predict_full(big)
            p, l = eng.predict_rows(big, np.float64)
            assert (out["proba1"] == p).all() and (out["label"] == l).all()
            s_alone = _alone_scores(iforest, thr, big)
            assert (out["outlier_score"] == s_alone[0]).all() and (out["is_outlier"] == s_alone[1]).all()
            # edge rows: numerics exactly on isolation-tree thresholds
            e = eng.predict_full(enc.encode_frame(iforest_edges))
            es = -iforest.decision_function(_num(iforest_edges))
            assert (e["is_outlier"] == (es > thr)).all() and np.abs(e["outlier_score"] - es).max() <= TOL32
        finally:
            eng.close()


def _alone_scores(iforest, thr, rows):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine

    eng = ForestEngine(flatten.flatten_isolation_forest(iforest, 9, 14, threshold=thr), 0)
    try:
        return eng.predict_rows(rows, np.float32)
    finally:
        eng.close()


def test_model_predict_outliers(curated, inference, iforest, rf100d6, tmp_path):
    """Plugin level: ``B200Model.predict`` returns the detector's flags; the reference threshold 0.95 never fires;
    NaN numerics are refused as the reference's pinned sklearn does; the artefact directory round-trips."""
    import joblib

    from oracle import reference_pipeline as rp

    from databricks_kubernetes_mlops_poc_b200 import load_model
    from databricks_kubernetes_mlops_poc_b200.model import B200Model, save_model_dir

    df = curated[rp.FEATURES].iloc[:5000]
    want_p, _ = rp.oracle_predict(rf100d6, df)
    want_s = -iforest.decision_function(_num(df))
    m = B200Model.from_pipeline(rf100d6, outlier=SimpleNamespace(isolationforest=iforest, threshold=0.95))
    try:
        out = m.predict(df)
        assert np.abs(np.asarray(out["predictions"]) - want_p).max() <= TOL64
        assert out["outliers"] == [0] * len(df)
    finally:
        m.close()
    m = B200Model.from_pipeline(rf100d6, outlier=iforest, outlier_threshold=0.0)
    try:
        for frame in (df, df.iloc[:1], df.iloc[:100], inference):
            out = m.predict(frame)
            ws = -iforest.decision_function(_num(frame))
            assert out["outliers"] == (ws > 0.0).astype(int).tolist()
            assert np.abs(np.asarray(out["predictions"]) - rp.oracle_predict(rf100d6, frame)[0]).max() <= TOL64
        assert 0 < sum(m.predict(df)["outliers"]) < len(df)
        bad = df.iloc[:10].copy()
        bad.iloc[3, bad.columns.get_loc(rp.NUMERIC_FEATURES[2])] = np.nan
        with pytest.raises(ValueError, match="NaN"):
            m.predict(bad)
        assert len(m.predict_proba1(bad)) == 10  # the classifier alone still imputes the median
        proba, flags = m.replicas[0].score(df)  # the server's per-GPU scoring path
        assert np.abs(proba - want_p).max() <= TOL64 and (flags == (want_s > 0.0)).all()
        save_model_dir(str(tmp_path / "a"), m.flat, outlier_blob=m.outlier_blob)
    finally:
        m.close()
    m2 = load_model(str(tmp_path / "a"))
    try:
        assert m2.predict(df)["outliers"] == (want_s > 0.0).astype(int).tolist()
    finally:
        m2.close()
    # the reference's own layout: artifacts/outlier.pkl next to artifacts/classifier/model/model.pkl
    d = tmp_path / "b" / "artifacts" / "classifier" / "model"
    d.mkdir(parents=True)
    joblib.dump(rf100d6, d / "model.pkl")
    joblib.dump(SimpleNamespace(isolationforest=iforest, threshold=0.01), tmp_path / "b" / "artifacts" / "outlier.pkl")
    m3 = load_model(str(tmp_path / "b"))
    try:
        assert m3.predict(df)["outliers"] == (want_s > 0.01).astype(int).tolist()
        assert (tmp_path / "b" / "outlier.b2f").exists()
    finally:
        m3.close()


def test_isolation_forest_against_frozen_library_outputs(curated, inference, iforest, rf100d6):
    """The GPU walk against tests/golden/expected_detectors.npz (sklearn scores frozen by make_golden_detectors.py);
    meaningful only when this box's sklearn grows the same isolation trees as the one that froze them."""
    import sklearn

    from oracle import datasets

    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine

    exp = datasets.load_expected("detectors")
    if str(exp["sklearn_version"]) != sklearn.__version__:
        pytest.skip("frozen with another sklearn version")
    head = curated.iloc[:3000]
    assert np.abs(-iforest.decision_function(_num(head)) - exp["iforest_score_head3000"]).max() <= 1e-12
    enc = RowEncoder(flatten.flatten_pipeline(rf100d6))
    eng = ForestEngine(flatten.flatten_isolation_forest(iforest, 9, 14, threshold=0.0), 0)
    try:
        for df, key in ((head, "iforest_score_head3000"), (inference, "iforest_score_inference")):
            s64, f64 = eng.predict_rows(enc.encode_frame(df), np.float64)
            assert np.abs(s64 - exp[key]).max() <= TOL64 and (f64 == (exp[key] > 0.0)).all()
    finally:
        eng.close()
