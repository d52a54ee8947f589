"""numpy emulation of the predict kernel's semantics over a forest blob (TEST INFRASTRUCTURE).

Mirrors ``csrc/forest_predict.cuh`` word for word -- impute, sentinel word, fixed-depth walk with
self-looping leaves, per-lane float64 partial sums, butterfly order -- so that the flattener and the
blob format can be checked against the oracle on a CPU-only box.  It is not a fallback: nothing in
the product imports it.
"""

import numpy as np

from databricks_kubernetes_mlops_poc_b200.flatten import META_CAT, META_FEAT_SHIFT, META_SLOT_MASK, SENTINEL_BITS, SENTINEL_WORD, parse_header


def walk_blob(blob: bytes, rows: np.ndarray):
    h = parse_header(blob)
    n = rows.shape[0]
    n_cat, n_num = h["n_cat"], h["n_num"]
    w = rows.astype(np.uint32).copy()
    f = w.view(np.float32)
    nan = np.isnan(f[:, n_cat : n_cat + n_num])
    imp = np.broadcast_to(h["impute"][n_cat : n_cat + n_num], nan.shape)
    f[:, n_cat : n_cat + n_num][nan] = imp[nan]
    w[:, SENTINEL_WORD] = SENTINEL_BITS
    lane_acc = np.zeros((n, 32), dtype=np.float64)
    buf = np.frombuffer(blob, dtype=np.uint8)
    ridx = np.arange(n)
    for g in h["groups"]:
        base = h["chunks_off"] + g["chunk_off"]
        ns, nl = g["n_slots"], g["n_leaf_slots"]
        N = buf[base : base + ns * 256].view(np.uint32).reshape(ns, 32, 2)
        T, M = N[:, :, 0], N[:, :, 1]
        LV = buf[base + ns * 256 : base + ns * 256 + nl * 256].view(np.float64).reshape(nl, 32)
        for lane in range(32):
            node = np.zeros(n, dtype=np.int64)
            for _ in range(g["depth"]):
                m = M[node, lane]
                t = T[node, lane]
                x = w[ridx, (m >> META_FEAT_SHIFT).astype(np.int64)]
                is_cat = (m & META_CAT) != 0
                with np.errstate(invalid="ignore"):
                    geu = ~(x.view(np.float32) < t.view(np.float32))  # x >= t or unordered
                second = (x == t) | (geu & ~is_cat)
                node = (m & np.uint32(META_SLOT_MASK)).astype(np.int64) + second.astype(np.int64)
            leaf = T[node, lane].astype(np.int64)
            lane_acc[:, lane] += LV[leaf, lane]
    v = lane_acc
    for o in (16, 8, 4, 2, 1):  # xor butterfly, as warp_sum()
        v = v + v[:, np.arange(32) ^ o]
    s = v[:, 0]
    if h["agg_mode"] == 0:
        return s / h["denom"], (s > (h["denom"] - s)).astype(np.int32)
    if h["agg_mode"] == 1:
        raw = h["init_raw"] + s
        return 1.0 / (1.0 + np.exp(-raw)), (raw >= 0).astype(np.int32)
    score = np.exp2(-(s / h["denom"])) + h["init_raw"]  # isolation forest (aggregate() in forest_predict.cuh)
    return score, (score > h["threshold"]).astype(np.int32)
