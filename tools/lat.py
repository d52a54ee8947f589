import sys, os, time, numpy as np
sys.path.insert(0, "/root/repo")
from databricks_kubernetes_mlops_poc_b200 import training, flatten
from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine
NT, MD = int(os.environ.get("NT", "500")), int(os.environ.get("MD", "8"))
base = training.load_base_frame()
pipe = training.fit_synthetic("rf", base, 20000, 1, n_estimators=NT, max_depth=MD, random_state=0)
flat = flatten.flatten_pipeline(pipe); enc = RowEncoder(flat); eng = ForestEngine(flat, 0)
print(eng.info())
N = 65536
_, codes, nums = training.synth_arrays(base, N, 3)
rows, proba, label = eng.staging(N)
enc.encode_arrays(codes, nums, out=rows)
pk = eng.pinned("pk", N*64).view(np.uint32, (N,16)); enc.pack_rows(rows, out=pk)
from databricks_kubernetes_mlops_poc_b200._cabi import SCORED_DTYPE
out = eng.pinned("out", N*8).view(SCORED_DTYPE, (N,))
d_rows = eng.device_alloc(rows.nbytes); d_p = eng.device_alloc(N*4); d_l = eng.device_alloc(N*4); eng.h2d(d_rows, rows)
for n in (1, 16, 256, 4096, 65536):
    ms = eng.predict_device_timed(d_rows, n, d_p, False, d_l, 50, False)
    for _ in range(20): eng.predict_pairs(pk[:n], out=out[:n])
    ts = []
    for _ in range(500 if n <= 4096 else 100):
        t0 = time.perf_counter(); eng.predict_pairs(pk[:n], out=out[:n]); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"n={n:6d} kernel med {np.median(ms)*1e3:7.2f} us | C-ABI call p50 {np.percentile(ts,50):7.1f} us p99 {np.percentile(ts,99):7.1f} us", flush=True)
