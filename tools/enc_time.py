"""Single-thread cost of the request encoder by part (65 536-row DataFrame): ranked / packed / 96-byte rows, category codes alone,
rank step alone.  CPU only."""
import sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from databricks_kubernetes_mlops_poc_b200 import flatten, training, _cabi
from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES
dist = bench.Dist(1, False, solo=True)
pipe, base = bench.get_pipeline("gbdt100d6", dist)
flat = flatten.flatten_pipeline(pipe)
enc = RowEncoder(flat)
pv, pc, pn = training.synth_arrays(base, 65536, 7)
df = training.arrays_to_frame(pv, pc, pn)[ALL_FEATURES]
print(df.dtypes.iloc[0], df.dtypes.iloc[10])
def best(f, k=7):
    t=[]
    for _ in range(k):
        t0=time.perf_counter(); f(); t.append(time.perf_counter()-t0)
    return min(t)*1e3
out = np.zeros((65536, enc.ranked_row_words), dtype=np.uint32)
assert enc._encode_native(df, out, fmt=2)
scol, ptrs, strides, keep = enc.frame_columns(df)
h = enc._native_handle(); L = enc._lib
for fmt in (2,1,0):
    o = np.zeros((65536, {2:enc.ranked_row_words,1:16,0:24}[fmt]), dtype=np.uint32)
    print("fmt",fmt,"encode t1 ms", best(lambda: L.b2f_encoder_encode(h, 65536, scol, ptrs, _cabi.ptr(strides), fmt, _cabi.ptr(o), 1)))
codes = np.zeros((9,65536), dtype=np.int32)
print("codes only t1 ms", best(lambda: L.b2f_encoder_codes(h, 65536, scol, _cabi.ptr(codes), 1)))
rows24 = enc.encode_frame(df)
print("rank_rows (from rows24) ms", best(lambda: enc.rank_rows(rows24)))
