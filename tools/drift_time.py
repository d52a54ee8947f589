"""Device time of one drift-scoring request (K3) by batch size, with scipy's own time beside it (gpurun helper)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from databricks_kubernetes_mlops_poc_b200.drift import TabularDrift  # noqa: E402
from oracle import datasets, drift as od, reference_pipeline as rp  # noqa: E402

cur = datasets.load_curated()
ref = cur[rp.FEATURES]
det = TabularDrift(ref, rp.CATEGORICAL_FEATURES, device=0)
rng = np.random.default_rng(0)
out = []
for n in (1, 16, 256, 1000, 4096, 65536):
    batch = ref.iloc[rng.integers(0, len(ref), n)].reset_index(drop=True)
    det.statistics(batch)
    dev, wall = [], []
    for _ in range(10 if n < 65536 else 3):
        t = time.perf_counter()
        det.statistics(batch)
        wall.append(time.perf_counter() - t)
        dev.append(det.last_device_ms)
    cpu = None
    if n <= 4096:
        t = time.perf_counter()
        od.tabular_drift_p_values(ref, batch, rp.CATEGORICAL_FEATURES)
        cpu = time.perf_counter() - t
    out.append(dict(n=n, device_ms=float(np.median(dev)), wall_ms=1e3 * float(np.median(wall)), scipy_ms=None if cpu is None else 1e3 * cpu))
    print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/drift_time.json", "w"), indent=1)
