"""The oracle itself, pinned (CPU): sklearn refit == frozen golden outputs; numpy and C restatements ==
sklearn; edge semantics the reference pipeline has (unknown / missing categories, NaN, inf)."""

import numpy as np
import pytest
import sklearn


def test_golden_frames_are_intact(curated, inference):
    assert curated.shape == (30000, 24), curated.shape
    assert inference.shape[1] == 23 and len(inference) >= 80, inference.shape
    assert list(inference.columns)[0] == "credit_limit"  # inference.csv column order differs from the model's
    assert abs(curated["default_payment_next_month"].mean() - 0.2212) < 1e-4


@pytest.mark.parametrize("name", ["rf100d6", "rf500d8"])
def test_refit_reproduces_frozen_outputs(curated, inference, name, request):
    """The pinned reference pipeline re-fit here gives the outputs frozen by tests/golden/make_golden.py
    (same sklearn version): labels identical, probabilities to summation-order noise."""
    from oracle import datasets
    from oracle import reference_pipeline as rp

    exp = datasets.load_expected(name)
    if str(exp["sklearn_version"]) != sklearn.__version__:
        pytest.skip(f"golden made with sklearn {exp['sklearn_version']}, running {sklearn.__version__}")
    pipe = request.getfixturevalue(name)
    p, l = rp.oracle_predict(pipe, curated)
    assert np.abs(p - exp["proba1"]).max() < 1e-13 and (l == exp["label"]).all()
    pi, li = rp.oracle_predict(pipe, inference)
    assert np.abs(pi - exp["inf_proba1"]).max() < 1e-13 and (li == exp["inf_label"]).all()
    assert int(exp["total_nodes"]) == sum(e.tree_.node_count for e in pipe.named_steps["classifier"].estimators_)
    assert float(exp["min_margin"]) > 1e-9  # no pinned row sits on the label knife edge


def test_restatements_match_library(curated, adversarial, rf100d6, gbdt_small):
    from oracle import reference_pipeline as rp
    from oracle import treewalk as tw

    for pipe in (rf100d6, gbdt_small):
        dump = tw.dump_pipeline(pipe)
        for df in (curated.iloc[:4000], adversarial):
            want_p, want_l = rp.oracle_predict(pipe, df)
            p, l = tw.predict_numpy(dump, df)
            assert np.abs(p - want_p).max() < 1e-14 and (l == want_l).all()
            codes, nums = tw.encode_frame(dump, df)
            pc, lc = tw.predict_c(dump, codes, nums)
            assert np.abs(pc - want_p).max() < 1e-14 and (lc == want_l).all()
            pc1, _ = tw.predict_c(dump, codes, nums, threads=1)
            assert (pc1 == pc).all()


def test_reference_edge_semantics(curated, rf100d6):
    """What the reference pipeline does at its edges -- these are the behaviours the GPU path must copy."""
    from oracle import reference_pipeline as rp
    from oracle import treewalk as tw

    df = curated[rp.FEATURES].iloc[:4].copy()
    base = rf100d6.predict_proba(df)[:, 1]
    unk = df.copy()
    unk["education"] = "never_seen"
    none = df.copy()
    none["education"] = None
    # unknown string and missing value both become the all-zero one-hot block
    assert np.allclose(rf100d6.predict_proba(unk)[:, 1], rf100d6.predict_proba(none)[:, 1], atol=0)
    nan = df.copy()
    nan["age"] = np.nan
    med = df.copy()
    med["age"] = np.median(rp.reference_split(curated)[0]["age"])
    assert np.allclose(rf100d6.predict_proba(nan)[:, 1], rf100d6.predict_proba(med)[:, 1], atol=0)
    inf = df.copy()
    inf["age"] = np.inf
    with pytest.raises(ValueError):
        rf100d6.predict_proba(inf)
    big = df.copy()
    big["age"] = 1e39  # finite in float64, overflows float32
    with pytest.raises(ValueError):
        rf100d6.predict_proba(big)
    dump = tw.dump_pipeline(rf100d6)
    with pytest.raises(ValueError):
        tw.predict_numpy(dump, big)
    codes, nums = tw.encode_frame(dump, big)
    with pytest.raises(ValueError):
        tw.predict_c(dump, codes, nums)
    assert base.shape == (4,)


def test_custom_model_restatement(curated, rf100d6):
    from oracle.custom_model import ReferenceCustomModel

    m = ReferenceCustomModel(rf100d6, curated)
    out = m.predict(None, curated.iloc[:50].drop(columns=["default_payment_next_month"]))
    assert set(out) == {"predictions", "outliers", "feature_drift_batch"}
    assert len(out["predictions"]) == 50 and out["outliers"] == [0] * 50
    assert len(out["feature_drift_batch"]) == 23
    assert all(0.0 <= v <= 1.0 for v in out["feature_drift_batch"].values())
