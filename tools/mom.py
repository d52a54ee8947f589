import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
from databricks_kubernetes_mlops_poc_b200 import training, flatten
from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine
base = training.load_base_frame()
pipe = training.fit_synthetic("rf", base, 3000, 1, n_estimators=10, max_depth=4, random_state=0)
flat = flatten.flatten_pipeline(pipe); enc = RowEncoder(flat); eng = ForestEngine(flat, 0)
N = 4 * 1024 * 1024
_, codes, nums = training.synth_arrays(base, N, 3)
rows = enc.encode_arrays(codes, nums)
d_rows = eng.device_alloc(rows.nbytes); eng.h2d(d_rows, rows)
for n in (256, 4096, 37888, 125000, 1000000, N):
    ms, out = eng.moments_device_timed(d_rows, n, 20, False)
    f = rows[:n].view(np.float32)[:, 9:23].astype(np.float64)
    ok = np.allclose(out[9:23, 1], np.nanmean(f, 0), rtol=1e-9) and np.allclose(out[9:23, 2] / out[9:23, 0], np.nanvar(f, 0), rtol=1e-8)
    print(n, "med %.2f us" % (np.median(ms) * 1e3), "GB/s %.0f" % (92 * n / np.median(ms) / 1e6), "frac %.3f" % (92 * n / np.median(ms) / 1e6 / 6568.4), "ok", ok, flush=True)
