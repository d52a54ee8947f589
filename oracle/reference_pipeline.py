"""Reference model definition, re-stated for the oracle (TEST INFRASTRUCTURE).

Follows the reference notebooks; nothing here is used by the product path.

* feature lists / column order ........ ``databricks/src/01-train-model.ipynb:126-158``
  (same lists at ``02-register-model.ipynb:119-151``; same order as the
  ``LoanApplicant`` fields, ``app/model.py:12-34``)
* pipeline definition .................. ``01-train-model.ipynb:195-231``
* train/test split ..................... ``01-train-model.ipynb:260-264``
* hyper-parameter domain ............... ``01-train-model.ipynb:343-345``

The reference's classifier is a RandomForest.  BASELINE.json's configs 2-4
name a "GBDT" model that the reference does not contain; for those the same
preprocessing is put in front of sklearn's ``GradientBoostingClassifier``
 -- that part has no reference counterpart and says so.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
from sklearn.compose import ColumnTransformer
from sklearn.ensemble import GradientBoostingClassifier, RandomForestClassifier
from sklearn.impute import SimpleImputer
from sklearn.model_selection import train_test_split
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import OneHotEncoder

TARGET = "default_payment_next_month"

CATEGORICAL_FEATURES = [
    "sex",
    "education",
    "marriage",
    "repayment_status_1",
    "repayment_status_2",
    "repayment_status_3",
    "repayment_status_4",
    "repayment_status_5",
    "repayment_status_6",
]

NUMERIC_FEATURES = [
    "credit_limit",
    "age",
    "bill_amount_1",
    "bill_amount_2",
    "bill_amount_3",
    "bill_amount_4",
    "bill_amount_5",
    "bill_amount_6",
    "payment_amount_1",
    "payment_amount_2",
    "payment_amount_3",
    "payment_amount_4",
    "payment_amount_5",
    "payment_amount_6",
]

FEATURES = CATEGORICAL_FEATURES + NUMERIC_FEATURES

# The two pinned reference models (inside the reference's search space,
# 01-train-model.ipynb:343-345) -- SURVEY.md section 8c.
PINNED_RF = {
    "rf100d6": dict(n_estimators=100, max_depth=6, criterion="gini", random_state=0),
    "rf500d8": dict(n_estimators=500, max_depth=8, criterion="entropy", random_state=0),
}
# GBDT stand-ins for BASELINE configs 2-4 (no reference counterpart).
PINNED_GBDT = {
    "gbdt100d6": dict(n_estimators=100, max_depth=6, random_state=0),
    "gbdt500d8": dict(n_estimators=500, max_depth=8, random_state=0),
}


def _preprocessor() -> ColumnTransformer:
    # 01-train-model.ipynb:197-221
    categorical_transformer = Pipeline(
        steps=[
            ("imputer", SimpleImputer(strategy="constant", fill_value="missing")),
            ("ohe", OneHotEncoder(handle_unknown="ignore")),
        ]
    )
    numeric_transformer = Pipeline(steps=[("imputer", SimpleImputer(strategy="median"))])
    return ColumnTransformer(
        transformers=[
            ("categorical", categorical_transformer, CATEGORICAL_FEATURES),
            ("numeric", numeric_transformer, NUMERIC_FEATURES),
        ]
    )


def make_classifier_pipeline(params: dict) -> Pipeline:
    """The reference's ``make_classifer_pipeline`` (01-train-model.ipynb:195-231)."""
    return Pipeline(
        [
            ("preprocessor", _preprocessor()),
            ("classifier", RandomForestClassifier(**params, n_jobs=-1)),
        ]
    )


def make_gbdt_pipeline(params: dict) -> Pipeline:
    """Same preprocessing, GBDT classifier (BASELINE configs 2-4; not in the reference)."""
    return Pipeline(
        [
            ("preprocessor", _preprocessor()),
            ("classifier", GradientBoostingClassifier(**params)),
        ]
    )


def reference_split(df: pd.DataFrame):
    """The reference's split (01-train-model.ipynb:260-264)."""
    return train_test_split(df[FEATURES + [TARGET]], test_size=0.20, random_state=2024)


def fit_reference_pipeline(df: pd.DataFrame, params: dict) -> Pipeline:
    """Fit as the reference does (01-train-model.ipynb:266-287) on its 80 % split."""
    df_train, _ = reference_split(df)
    est = make_classifier_pipeline(params)
    est.fit(df_train[FEATURES], df_train[[TARGET]].values.ravel())
    return est


def fit_gbdt_pipeline(df_train: pd.DataFrame, y: np.ndarray, params: dict) -> Pipeline:
    est = make_gbdt_pipeline(params)
    est.fit(df_train[FEATURES], np.asarray(y).ravel())
    return est


def oracle_predict(pipeline: Pipeline, df: pd.DataFrame):
    """What the reference's serving path computes with the classifier:
    ``classifier.predict_proba(df[all_features])[:, 1]`` (02-register-model.ipynb:335-337)
    plus the hard label ``predict`` (used at train time, 01-train-model.ipynb:290)."""
    X = df[FEATURES]
    proba1 = pipeline.predict_proba(X)[:, 1]
    label = pipeline.predict(X).astype(np.int32)
    return proba1, label
