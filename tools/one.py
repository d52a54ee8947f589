import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
from databricks_kubernetes_mlops_poc_b200 import training, flatten
from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine
base = training.load_base_frame()
pipe = training.fit_synthetic("rf", base, 5000, 1, n_estimators=100, max_depth=6, random_state=0)
flat = flatten.flatten_pipeline(pipe); enc = RowEncoder(flat); eng = ForestEngine(flat, 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
_, codes, nums = training.synth_arrays(base, N, 3)
rows = enc.encode_arrays(codes, nums)
d_rows = eng.device_alloc(rows.nbytes); d_p = eng.device_alloc(N*4); d_l = eng.device_alloc(N*4)
eng.h2d(d_rows, rows)
ms = eng.predict_device_timed(d_rows, N, d_p, False, d_l, 5, False)
print(N, ms)
