import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
from databricks_kubernetes_mlops_poc_b200 import training, flatten
from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine
base = training.load_base_frame()
pipe = training.fit_synthetic("rf", base, 5000, 1, n_estimators=100, max_depth=6, random_state=0)
flat = flatten.flatten_pipeline(pipe); enc = RowEncoder(flat); eng = ForestEngine(flat, 0)
n = 65536
_, codes, nums = training.synth_arrays(base, n, 3)
rows, proba, label = eng.staging(n)
enc.encode_arrays(codes, nums, out=rows)
p32 = proba.view(np.float32)[:n]
for i in range(6):
    eng.predict_rows(rows, np.float32, out_proba=p32, out_label=label)
