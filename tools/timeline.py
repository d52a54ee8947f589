import sys, os, time, numpy as np
sys.path.insert(0, "/root/repo")
from databricks_kubernetes_mlops_poc_b200 import training, flatten
from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine
from databricks_kubernetes_mlops_poc_b200._cabi import SCORED_DTYPE
base = training.load_base_frame()
pipe = training.fit_synthetic("rf", base, 5000, 1, n_estimators=100, max_depth=6, random_state=0)
flat = flatten.flatten_pipeline(pipe); enc = RowEncoder(flat); eng = ForestEngine(flat, 0)
n = 65536
_, codes, nums = training.synth_arrays(base, n, 3)
pk = eng.pinned("pk", n*64).view(np.uint32, (n,16)); enc.encode_arrays_packed(codes, nums, out=pk)
out = eng.pinned("out", n*8).view(SCORED_DTYPE, (n,))
for i in range(8):
    t0=time.perf_counter(); eng.predict_pairs(pk, out=out); print("call us", (time.perf_counter()-t0)*1e6, file=sys.stderr)
