"""Ranked rows and the rank layout of the forest, checked on a CPU-only box (no GPU involved).

* the host-side ranking (csrc/forest_rank.h + csrc/host_simd.cpp, every SIMD level) against a brute-force
  ``searchsorted`` over the forest's own split values;
* the rank layout (complete trees of 4-byte nodes, every test a 16-bit rank test) walked by the numpy emulator
  ``tests/rank_walk.py`` against the library (sklearn) on reference rows and on the adversarial rows -- the same
  oracle the GPU kernel is held to."""

import os
import subprocess
import sys

import numpy as np
import pytest

from rank_walk import unpack_ranked, walk_rank_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _enc(pipe):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder

    flat = flatten.flatten_pipeline(pipe)
    return flat, RowEncoder(flat)


def test_rank_info_and_row_layout(rf100d6):
    flat, enc = _enc(rf100d6)
    info = enc.rank_info()
    assert info.ok and info.depth == 6 and info.n_trees == 100
    assert info.cat_bytes == 4 and info.row_bytes == 32  # 9 categorical fields in 32 bits + 14 uint16 ranks
    assert info.layout_bytes == 104 * (1 << 6) * 12  # trees padded to a multiple of 8, 2^D * (4 + 8) bytes each
    assert info.n_pairs <= sum(len(c) for c in flat.categories)
    for k in range(14):
        thr = enc.rank_thresholds(k)
        assert len(thr) == info.n_thresholds[k] and (np.diff(thr) > 0).all()


def test_ranks_equal_brute_force(curated, adversarial, rf100d6):
    from databricks_kubernetes_mlops_poc_b200.flatten import parse_header

    flat, enc = _enc(rf100d6)
    info = enc.rank_info()
    h = parse_header(flat.blob)
    for df in (curated.iloc[:5000], adversarial):
        rows = enc.encode_frame(df)
        rk = enc.rank_rows(rows, threads=3)
        vals = unpack_ranked(rk, info)
        x = rows.view(np.float32)[:, 9:23].copy()
        for k in range(14):
            col = x[:, k].copy()
            col[np.isnan(col)] = h["impute"][9 + k]
            want = np.searchsorted(enc.rank_thresholds(k), col, side="right")
            assert (vals[:, k] == want).all(), f"numeric {k}"
        # packed rows rank to the same thing
        assert np.array_equal(enc.rank_rows(enc.pack_rows(rows)), rk)


@pytest.mark.parametrize("level", ["0", "1", "2"])
def test_simd_levels_agree(level):
    """Scalar, AVX2 and AVX-512 forms of the rank-table walk give identical rows (B2F_SIMD caps the level)."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from oracle import datasets, reference_pipeline as rp\n"
        "from databricks_kubernetes_mlops_poc_b200 import flatten\n"
        "from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder\n"
        "cur = datasets.load_curated()\n"
        "pipe = rp.fit_reference_pipeline(cur.iloc[:3000], dict(n_estimators=20, max_depth=6, random_state=0))\n"
        "enc = RowEncoder(flatten.flatten_pipeline(pipe))\n"
        "rk = enc.rank_rows(enc.encode_frame(cur.iloc[3000:6000]))\n"
        "import hashlib; print(hashlib.sha256(rk.tobytes()).hexdigest())\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, B2F_SIMD=level)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    digest = out.stdout.strip().splitlines()[-1]
    ref = getattr(test_simd_levels_agree, "_ref", None)
    if ref is None:
        test_simd_levels_agree._ref = digest
    else:
        assert digest == ref


@pytest.mark.parametrize("which", ["rf100d6", "gbdt_small"])
def test_rank_layout_walk_matches_library(curated, inference, adversarial, rf100d6, gbdt_small, which):
    from oracle import reference_pipeline as rp

    pipe = {"rf100d6": rf100d6, "gbdt_small": gbdt_small}[which]
    flat, enc = _enc(pipe)
    info = enc.rank_info()
    assert info.ok
    layout = enc.rank_layout()
    for df in (curated.iloc[:3000], inference, adversarial):
        want_p, want_l = rp.oracle_predict(pipe, df)
        p, l = walk_rank_layout(layout, info, flat.blob, enc.rank_rows(enc.encode_frame(df)))
        assert np.abs(p - want_p).max() <= 1e-12 and (l == want_l).all()
        if len(df) > 128:  # the native encoder writes the same ranked rows straight from the DataFrame's buffers
            assert np.array_equal(enc.encode_frame_ranked(df), enc.rank_rows(enc.encode_frame(df)))


def test_forests_without_a_rank_layout(curated):
    """Trees deeper than 8 levels keep the float32 kernels; the ranker says why."""
    from oracle import reference_pipeline as rp

    pipe = rp.fit_reference_pipeline(curated.iloc[:3000], dict(n_estimators=5, max_depth=12, random_state=0))
    flat, enc = _enc(pipe)
    info = enc.rank_info()
    assert not info.ok and b"deeper" in info.why
    with pytest.raises(ValueError):
        enc.rank_rows(enc.encode_frame(curated.iloc[:10]))
