"""Restatement of the reference's ``CustomModel`` (TEST INFRASTRUCTURE).

Follows ``databricks/src/02-register-model.ipynb:305-353``: ``predict`` builds a DataFrame, takes
``classifier.predict_proba(df[all_features])[:, 1]``, the drift detector's p-values and the outlier
detector's flags, and returns the response dict.  mlflow and alibi-detect are not installed in this
image: the classifier part is exact (the real sklearn pipeline); the drift part
restates alibi-detect 0.12.0's ``TabularDrift`` with scipy (``oracle/drift.py``: chi-squared on category counts
over the union of reference and batch categories, exact two-sample K-S on numerics, float32 p-values) and is
therefore UNPINNED against the real package; the
outlier part uses sklearn's IsolationForest the way alibi-detect's ``IForest`` does
(``score = -decision_function``, ``is_outlier = score > threshold``) with the reference's threshold 0.95,
which can never fire (the score is bounded by 0.5), so the flags are all 0.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
from sklearn.ensemble import IsolationForest

from .reference_pipeline import CATEGORICAL_FEATURES, FEATURES, NUMERIC_FEATURES


class ReferenceCustomModel:
    def __init__(self, classifier, reference_frame: pd.DataFrame, outlier_trees: int = 100):
        # 02-register-model.ipynb:310-315
        self.categorical_features = list(CATEGORICAL_FEATURES)
        self.numeric_features = list(NUMERIC_FEATURES)
        self.all_features = list(FEATURES)
        self.classifier = classifier
        # 02-register-model.ipynb:224-229: TabularDrift(x_ref, p_val=.05, categories_per_feature={0..8: None})
        self.x_ref = reference_frame[self.all_features]
        # 02-register-model.ipynb:232-233: IForest(threshold=0.95).fit(df[NUMERIC_FEATURES].values)
        self.threshold = 0.95
        self.iforest = IsolationForest(n_estimators=outlier_trees, random_state=0).fit(
            reference_frame[self.numeric_features].to_numpy()
        )

    def drift_p_values(self, df: pd.DataFrame) -> np.ndarray:
        from .drift import tabular_drift_p_values

        return tabular_drift_p_values(self.x_ref, df[self.all_features], self.categorical_features)

    def predict(self, context, model_input):
        df = pd.DataFrame(model_input)  # :332
        predictions = self.classifier.predict_proba(df[self.all_features])[:, 1].tolist()  # :335-337
        p_val = self.drift_p_values(df[self.all_features])  # :338
        score = -self.iforest.decision_function(df[self.numeric_features].to_numpy())  # :339
        is_outlier = (score > self.threshold).astype(int)
        return {  # :342-353
            "predictions": predictions,
            "outliers": is_outlier.tolist(),
            "feature_drift_batch": dict(zip(self.all_features, (1 - p_val).tolist())),
        }
