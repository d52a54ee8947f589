"""Launch the K3 (drift) or K4 (isolation forest) kernels a few times -- the command ncu wraps (gpurun helper)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import datasets, reference_pipeline as rp  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "drift"
cur = datasets.load_curated()
ref = cur[rp.FEATURES]
rng = np.random.default_rng(0)
if which == "drift":
    from databricks_kubernetes_mlops_poc_b200.drift import TabularDrift

    det = TabularDrift(ref, rp.CATEGORICAL_FEATURES, device=0)
    for n in (1, 1000):
        batch = ref.iloc[rng.integers(0, len(ref), n)].reset_index(drop=True)
        det.statistics(batch)
        print(n, det.last_device_ms)
    det.close()
else:
    from sklearn.ensemble import IsolationForest

    from databricks_kubernetes_mlops_poc_b200 import flatten, training
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine

    pipe = rp.fit_reference_pipeline(cur.iloc[:3000], dict(n_estimators=10, max_depth=4, random_state=0))
    enc = RowEncoder(flatten.flatten_pipeline(pipe))
    iso = IsolationForest(n_estimators=100, random_state=0).fit(ref[rp.NUMERIC_FEATURES].to_numpy())
    eng = ForestEngine(flatten.flatten_isolation_forest(iso, 9, 14, threshold=0.0), 0)
    _, codes, nums = training.synth_arrays(training.load_base_frame(), 65536, 3)
    rows = enc.encode_arrays_packed(codes, nums)
    for _ in range(2):
        s, f = eng.predict_rows(rows, np.float32)
    print(eng.info())
    eng.close()
