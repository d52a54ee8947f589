"""numpy emulation of k_forest_predict_rank over the forest's rank layout (TEST INFRASTRUCTURE).

Mirrors ``csrc/forest_predict_rank.cuh``: ranked rows are unpacked into the kernel's 16-bit values (numeric k: its rank;
tested (categorical feature, category) pair i: 0 / 1), every node is the test ``value[f] >= t`` evaluated as
``(value << 16 | 0xFFFF) >= node word``, every tree is walked as a complete binary tree (child of node i = 2i + 1 + second),
the path bits index the float64 payloads, trees are added in order.
It lets the rank layout built by ``csrc/forest_rank.h`` and the host-side ranking be checked against the oracle on a
CPU-only box.  It is not a fallback: nothing in the product imports it."""

import numpy as np

from databricks_kubernetes_mlops_poc_b200.flatten import parse_header


def unpack_ranked(rows: np.ndarray, info) -> np.ndarray:
    """ranked rows (N, row_bytes/4) uint32 -> the kernel's per-row 16-bit values (N, n_num + n_pairs) as phase 1 builds them:
    pseudo-feature k < n_num = rank of numeric k; even(n_num) + i = 1 iff the row's code of pair i's feature is pair i's category."""
    raw = np.ascontiguousarray(rows).view(np.uint8).reshape(rows.shape[0], -1)
    n = raw.shape[0]
    cw = np.zeros(n, dtype=np.uint64)
    for b in range(info.cat_bytes):
        cw |= raw[:, b].astype(np.uint64) << np.uint64(8 * b)
    base = (info.n_num + 1) & ~1  # one-hot values start at an even pseudo-feature index
    out = np.zeros((n, base + info.n_pairs), dtype=np.uint32)
    ranks = raw[:, info.cat_bytes : info.cat_bytes + 2 * info.n_num].copy().view(np.uint16).reshape(n, info.n_num)
    out[:, : info.n_num] = ranks
    for i in range(info.n_pairs):
        j, c = info.pairs[i] >> 16, info.pairs[i] & 0xFFFF
        code1 = (cw >> np.uint64(info.cat_shift[j])) & np.uint64((1 << info.cat_bits[j]) - 1)
        out[:, base + i] = code1 == np.uint64(c + 1)
    return out


def walk_rank_layout(layout: np.ndarray, info, blob: bytes, rows: np.ndarray):
    h = parse_header(blob)
    x = unpack_ranked(rows, info)
    n = x.shape[0]
    D = info.depth
    slots = 1 << D
    stride = slots * 12
    ridx = np.arange(n)
    s = np.zeros(n, dtype=np.float64) if h["agg_mode"] != 1 else np.full(n, h["init_raw"], dtype=np.float64)
    for t in range(info.n_trees):
        base = t * stride
        nodes = layout[base : base + slots * 4].view(np.uint32)
        leaves = layout[base + slots * 4 : base + slots * 12].view(np.float64)
        i = np.zeros(n, dtype=np.int64)
        for _ in range(D):
            nw = nodes[i]
            off = (nw & np.uint32(0xFFFF)).astype(np.int64)
            f = (off >> 7) * 2 + ((off >> 1) & 1)  # byte offset (f >> 1) * 128 + (f & 1) * 2 -> pseudo-feature
            v = x[ridx, np.minimum(f, x.shape[1] - 1)]
            second = ((v << np.uint32(16)) | np.uint32(0xFFFF)) >= nw
            i = 2 * i + 1 + second.astype(np.int64)
        s += leaves[i - (slots - 1)]
    if h["agg_mode"] == 0:
        return s / h["denom"], (s > (h["denom"] - s)).astype(np.int32)
    if h["agg_mode"] == 1:
        return 1.0 / (1.0 + np.exp(-s)), (s >= 0).astype(np.int32)
    score = np.exp2(-(s / h["denom"])) + h["init_raw"]
    return score, (score > h["threshold"]).astype(np.int32)
