"""HTTP surface with a stub scorer (CPU): status codes, schema, logging, cross-request batching.
Mirrors the behaviours probed on the unmodified reference app."""

import asyncio
import json
import logging

import numpy as np
import pytest
from fastapi.testclient import TestClient


class StubModel:
    """Deterministic scorer standing in for the GPU model: P = (credit_limit mod 1000) / 1000."""

    drift = None

    def __init__(self, fail=False):
        self.calls, self.fail = [], fail
        self.replicas = [self]

    def predict_proba1(self, df):
        if self.fail:
            raise RuntimeError("b2f_predict failed (rc=-2): CUDA error")
        self.calls.append(len(df))
        return (df["credit_limit"].to_numpy() % 1000) / 1000.0


def _client(model):
    from databricks_kubernetes_mlops_poc_b200.server import create_app

    return TestClient(create_app(model=model), raise_server_exceptions=False)


def test_predict_contract(caplog):
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES, sample_request

    m = StubModel()
    with caplog.at_level(logging.INFO), _client(m) as c:
        r = c.post("/predict", json=sample_request())  # the reference CI smoke test body
        assert r.status_code == 200
        body = r.json()
        assert set(body) == {"predictions", "outliers", "feature_drift_batch"}
        assert body["predictions"] == [0.0] and body["outliers"] == [0.0]  # int flags serialised as floats
        assert list(body["feature_drift_batch"]) == ALL_FEATURES
        r = c.post("/predict", json=[{"credit_limit": 1250.0}, {}, {"sex": "female", "credit_limit": 333}])
        assert r.status_code == 200 and r.json()["predictions"] == [0.25, 0.0, 0.333]
        assert c.post("/predict", json=[{}]).status_code == 200  # defaults make {} a valid row
        assert c.post("/predict", json=[{"sex": 3}]).status_code == 422
        assert c.post("/predict", json={"sex": "male"}).status_code == 422
        assert c.post("/predict", json=[]).status_code == 500  # as the reference (empty DataFrame)
        assert c.get("/").status_code == 200  # Swagger UI at the root
    import time

    time.sleep(0.2)  # log lines are produced off the request path
    recs = [json.loads(r.getMessage()) for r in caplog.records if r.getMessage().startswith("{")]
    kinds = [r["type"] for r in recs]
    assert "InferenceData" in kinds and "ModelOutput" in kinds
    inf = next(r for r in recs if r["type"] == "InferenceData")
    assert inf["service_name"] == "credit-default-api" and len(inf["request_id"]) == 32
    assert json.loads(inf["data"])[0]["sex"] == "male"  # data is a JSON *string* of records, as in the reference
    out = next(r for r in recs if r["type"] == "ModelOutput" and r["request_id"] == inf["request_id"])
    assert set(out["data"]) == {"predictions", "outliers", "feature_drift_batch"}


def test_outlier_flags_pass_through():
    """A replica with ``score`` (classifier + outlier forest in one pass) feeds the response's ``outliers``."""

    class Scoring(StubModel):
        def score(self, df):
            x = df["credit_limit"].to_numpy()
            return (x % 1000) / 1000.0, (x > 5000).astype(np.int32)

    with _client(Scoring()) as c:
        r = c.post("/predict", json=[{"credit_limit": 1250.0}, {"credit_limit": 9100.0}, {}])
        assert r.status_code == 200
        assert r.json()["predictions"] == [0.25, 0.1, 0.0] and r.json()["outliers"] == [0.0, 1.0, 1.0]  # {} = the schema defaults (credit_limit 18000)


def test_engine_failure_is_http_500():
    with _client(StubModel(fail=True)) as c:
        assert c.post("/predict", json=[{}]).status_code == 500


def test_concurrent_requests_share_batches():
    import pandas as pd

    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES, DEFAULTS
    from databricks_kubernetes_mlops_poc_b200.server import MicroBatcher

    m = StubModel()
    mb = MicroBatcher([m], max_rows=4096, window_us=20000)

    async def main():
        frames = [pd.DataFrame([{**DEFAULTS, "credit_limit": float(1000 * i + j)} for j in range(3)])[ALL_FEATURES] for i in range(40)]
        outs = await asyncio.gather(*[mb.score(f) for f in frames])
        for i, o in enumerate(outs):
            assert np.allclose(o[0], [0.0, 0.001, 0.002]) and o[1] is None
        return len(outs)

    try:
        assert asyncio.run(main()) == 40
    finally:
        mb.close()
    assert sum(m.calls) == 120 and len(m.calls) < 40  # requests were merged into fewer engine calls


REFERENCE_APP = "/root/reference/app"


@pytest.mark.skipif(not __import__("os").path.exists(REFERENCE_APP + "/main.py"), reason="reference checkout not present (GPU box)")
def test_unmodified_reference_app_runs_on_the_shim(monkeypatch):
    """the reference's own app/main.py, imported unmodified from /root/reference with the mlflow shim
    ahead on the path, serves POST /predict from whatever `load_model` returns (a stub scorer here: no GPU)."""
    import importlib
    import os
    import sys

    import databricks_kubernetes_mlops_poc_b200 as pkg
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES, sample_request

    class Plugin:  # the plugin boundary: predict(DataFrame) -> dict (CustomModel.predict)
        def predict(self, df):
            if len(df.columns) == 0:
                raise KeyError("no columns")  # what B200Model.predict (and the reference's CustomModel) do on []
            n = len(df)
            return {"predictions": [0.25] * n, "outliers": [0] * n, "feature_drift_batch": {k: 0.0 for k in ALL_FEATURES}}

    monkeypatch.syspath_prepend(REFERENCE_APP)
    monkeypatch.syspath_prepend(os.path.join(os.path.dirname(pkg.__file__), "shim"))
    for name in [m for m in sys.modules if m in ("mlflow", "main", "model") or m.startswith("mlflow.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.setattr(pkg, "load_model", lambda path: Plugin())
    main = importlib.import_module("main")
    try:
        assert main.__file__.startswith(REFERENCE_APP)
        with TestClient(main.app, raise_server_exceptions=False) as c:
            r = c.post("/predict", json=sample_request())
            assert r.status_code == 200 and r.json()["predictions"] == [0.25]
            assert list(r.json()["feature_drift_batch"]) == ALL_FEATURES
            assert c.post("/predict", json=[]).status_code == 500
    finally:
        for name in [m for m in sys.modules if m in ("mlflow", "main", "model") or m.startswith("mlflow.")]:
            sys.modules.pop(name, None)


def test_one_pass_request_parsing_equals_the_model_validation():
    """parse_request (pydantic-core over the raw bytes into dict rows) accepts, rejects and coerces exactly like
    FastAPI's `data: list[LoanApplicant]` (json.loads + one model per row), and the docs keep the request schema."""
    from fastapi.exceptions import RequestValidationError
    from pydantic import TypeAdapter, ValidationError

    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES, DEFAULTS, LoanApplicant
    from databricks_kubernetes_mlops_poc_b200.server import create_app, parse_request, rows_to_frame

    model_rows = TypeAdapter(list[LoanApplicant])
    good = [
        b"[]", b"[{}]", b'[{"sex": "female", "age": 41}]', b'[{"age": "41", "credit_limit": 5, "unknown_key": 1}]',
        b'[{"bill_amount_1": 1e3, "education": ""}, {"payment_amount_6": -0.0}]', b'[{"age": true}]',
    ]
    for raw in good:
        want = model_rows.validate_json(raw)
        got = parse_request(raw)
        assert len(got) == len(want)
        if want:
            a, b = rows_to_frame(got), rows_to_frame(want)
            assert list(a.columns) == ALL_FEATURES and a.equals(b)
    assert rows_to_frame(parse_request(b"[{}]")).iloc[0].to_dict() == DEFAULTS
    bad = [b"", b"{", b'{"sex": "male"}', b'[{"sex": 3}]', b'[{"age": "old"}]', b"[1]", b'[{"age": null}]', b'[{"sex": null}]', b"null"]
    for raw in bad:
        with pytest.raises(ValidationError):
            model_rows.validate_json(raw)
        with pytest.raises(RequestValidationError) as ei:
            parse_request(raw)
        assert all(err["loc"][0] == "body" for err in ei.value.errors())
    with _client(StubModel()) as c:
        for raw in bad:
            r = c.post("/predict", content=raw, headers={"content-type": "application/json"})
            assert r.status_code == 422 and "detail" in r.json()
        spec = c.get("/openapi.json").json()
        body = spec["paths"]["/predict"]["post"]["requestBody"]["content"]["application/json"]["schema"]
        assert body["type"] == "array" and list(body["items"]["properties"]) == ALL_FEATURES
        assert body["items"]["properties"]["age"]["default"] == 18000.0
