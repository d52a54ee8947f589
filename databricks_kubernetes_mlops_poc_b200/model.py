"""B200Model: the drop-in for the object ``mlflow.pyfunc.load_model(dir)`` returns in the reference.

Plugin boundary being mirrored:

* reference ``app/main.py:26-28``  ``ml_models["credit_default"] = mlflow.pyfunc.load_model(MODEL_DIRECTORY)``
* reference ``app/main.py:72``     ``model_output = ml_models["credit_default"].predict(input_df)``
* reference ``CustomModel.predict`` (``databricks/src/02-register-model.ipynb:330-353``): returns
  ``{"predictions": [...], "outliers": [...], "feature_drift_batch": {23 names -> float}}``.

``predictions`` (the accelerated path) comes from the CUDA engine -- dictionary-encode on
the host into pinned memory, H2D, fused kernel, D2H -- with no CPU fallback.  ``feature_drift_batch``
and ``outliers`` ("next" rows) complete the response schema: drift through the
optional CPU detector in ``drift.py``, outliers as the constant 0 the reference provably returns
(its ``IForest(threshold=0.95)`` compares a score bounded by 0.5 with 0.95).
"""

from __future__ import annotations

import os

import numpy as np
import pandas as pd

from .encode import RowEncoder
from .engine import EngineGroup, ForestEngine
from .flatten import FlatForest, flatten_pipeline

BLOB_FILE = "forest.b2f.npz"
DRIFT_FILE = "drift_reference.npz"
SKLEARN_PICKLE = os.path.join("artifacts", "classifier", "model", "model.pkl")  # MLflow layout, 02-register-model.ipynb:317-321


class B200Model:
    def __init__(self, flat: FlatForest, devices=None, drift=None, proba_dtype=np.float64):
        self.flat = flat
        self.all_features = flat.all_features
        self.categorical_features = list(flat.cat_features)
        self.numeric_features = list(flat.num_features)
        self.encoder = RowEncoder(flat)
        devices = [0] if devices is None else list(devices)
        if len(devices) == 1:
            self.engine = ForestEngine(flat, devices[0])
            self.group = None
        else:
            self.group = EngineGroup(flat, devices)
            self.engine = self.group.engines[0]
        self.drift = drift
        self.proba_dtype = np.dtype(proba_dtype)
        self.classes = np.asarray(flat.classes)
        # one scoring replica per GPU for the server's round-robin batcher (each has its own handle,
        # pinned staging and worker thread; the forest is replicated, rows are independent)
        engines = self.group.engines if self.group is not None else [self.engine]
        self.replicas = [_Replica(self.encoder, e) for e in engines]

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pipeline(cls, pipeline, reference_frame: pd.DataFrame | None = None, **kw) -> "B200Model":
        """Fitted sklearn Pipeline (the reference's model.pkl) -> model on the GPU."""
        flat = flatten_pipeline(pipeline)
        drift = None
        if reference_frame is not None:
            from .drift import TabularDriftCPU

            drift = TabularDriftCPU(reference_frame[flat.all_features], flat.cat_features)
        return cls(flat, drift=drift, **kw)

    def close(self) -> None:
        if self.group is not None:
            self.group.close()
        else:
            self.engine.close()

    # ------------------------------------------------------------------ scoring
    def _score(self, df: pd.DataFrame):
        n = len(df)
        # large requests travel as 64-byte packed rows (one third fewer PCIe bytes), encoded natively in one pass
        packed = self.encoder.packed_ok and n > self.encoder.SMALL_BATCH
        rows, proba, label = self.engine.staging(n, packed=packed)
        if packed:
            self.encoder.encode_frame_packed(df, out=rows)
        else:
            self.encoder.encode_frame(df, out=rows)
        if self.proba_dtype != np.float64:
            proba = proba.view(np.float32)[:n]
        target = self.group if self.group is not None else self.engine
        target.predict_rows(rows, proba_dtype=self.proba_dtype, out_proba=proba, out_label=label)
        return proba, label

    def predict_proba1(self, df: pd.DataFrame) -> np.ndarray:
        """``classifier.predict_proba(df[all_features])[:, 1]`` (02-register-model.ipynb:335-337)."""
        return np.array(self._score(df)[0], dtype=np.float64)

    def predict_label(self, df: pd.DataFrame) -> np.ndarray:
        """``pipeline.predict(df)`` (hard labels, 01-train-model.ipynb:290)."""
        return self.classes[np.array(self._score(df)[1])]

    def predict(self, model_input) -> dict:
        """Mirror of ``CustomModel.predict(context, model_input)`` (02-register-model.ipynb:330-353)."""
        df = pd.DataFrame(model_input)
        if len(df.columns) == 0:
            # the reference dies in df[self.all_features] on an empty request (-> HTTP 500)
            raise KeyError(f"None of {self.all_features} are in the [columns]")
        proba, _ = self._score(df)
        n = len(df)
        if self.drift is not None:
            drift_scores = self.drift.score(df[self.all_features])
        else:
            drift_scores = [0.0] * len(self.all_features)
        return {
            "predictions": proba.tolist(),
            "outliers": [0] * n,
            "feature_drift_batch": dict(zip(self.all_features, drift_scores)),
        }


class _Replica:
    """One GPU's view of the model: encode into that engine's pinned staging and score there."""

    def __init__(self, encoder: RowEncoder, engine: ForestEngine):
        self.encoder, self.engine = encoder, engine

    def predict_proba1(self, df: pd.DataFrame) -> np.ndarray:
        n = len(df)
        packed = self.encoder.packed_ok and n > self.encoder.SMALL_BATCH
        rows, proba, label = self.engine.staging(n, packed=packed)
        if packed:
            self.encoder.encode_frame_packed(df, out=rows)
        else:
            self.encoder.encode_frame(df, out=rows)
        self.engine.predict_rows(rows, proba_dtype=np.float64, want_label=False, out_proba=proba)
        return np.array(proba, dtype=np.float64)


# ---------------------------------------------------------------------- loading
def save_model_dir(path: str, flat: FlatForest, reference_frame: pd.DataFrame | None = None) -> None:
    """Write the GPU-side artefact next to (or instead of) the MLflow pickles."""
    os.makedirs(path, exist_ok=True)
    flat.save(os.path.join(path, BLOB_FILE))
    if reference_frame is not None:
        from .drift import TabularDriftCPU

        TabularDriftCPU(reference_frame[flat.all_features], flat.cat_features).save(os.path.join(path, DRIFT_FILE))


def load_model(path: str, devices=None, **kw) -> B200Model:
    """Drop-in for ``mlflow.pyfunc.load_model(path)`` as used at reference ``app/main.py:26-28``.

    Looks for the cached forest blob first; otherwise for the sklearn pipeline pickle in the MLflow
    artefact layout (only loadable when the pickle's sklearn version matches) and flattens it.
    """
    blob_path = os.path.join(path, BLOB_FILE)
    if os.path.exists(blob_path):
        flat = FlatForest.load(blob_path)
    else:
        pkl = os.path.join(path, SKLEARN_PICKLE)
        if not os.path.exists(pkl):
            raise FileNotFoundError(f"neither {blob_path} nor {pkl} exists")
        import joblib

        flat = flatten_pipeline(joblib.load(pkl))
        try:
            flat.save(blob_path)
        except OSError:
            pass  # read-only image: keep the blob in memory only
    drift = None
    drift_path = os.path.join(path, DRIFT_FILE)
    if os.path.exists(drift_path) and os.environ.get("B200_DRIFT", "cpu") != "off":
        from .drift import TabularDriftCPU

        drift = TabularDriftCPU.load(drift_path)
    if devices is None:
        env = os.environ.get("B200_DEVICES")
        devices = [int(d) for d in env.split(",")] if env else [0]
    return B200Model(flat, devices=devices, drift=drift, **kw)
