"""CPU tests of the round-2 host code: one bad request must fail alone in a shared batch, 422 bodies equal the reference's,
bounded log backlog, packed rows only for the 9-categorical layout, recycled float lists, block-manager column access."""

import asyncio
import gc

import numpy as np
import pandas as pd
import pytest


# ----------------------------------------------------------------------------- server
class PickyModel:
    """Scores like the stub, but refuses any frame that holds a credit_limit above float32 range -- what the row encoder
    does (sklearn raises ValueError there)."""

    drift = None

    def __init__(self):
        self.calls = []
        self.replicas = [self]

    def predict_proba1(self, df):
        self.calls.append(len(df))
        x = df["credit_limit"].to_numpy()
        if (np.abs(x) > 3.4e38).any():
            raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
        return (x % 1000) / 1000.0


def test_one_bad_request_fails_alone_in_a_shared_batch():
    """The reference scores requests independently (app/main.py:72): when the micro-batcher has merged several requests and
    the model rejects the merged frame, every request is re-scored on its own and only the offender gets the exception."""
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES, DEFAULTS
    from databricks_kubernetes_mlops_poc_b200.server import MicroBatcher

    m = PickyModel()
    mb = MicroBatcher([m], max_rows=4096, window_us=50000)

    async def main():
        frames = [pd.DataFrame([{**DEFAULTS, "credit_limit": float(1000 * i + 7)}])[ALL_FEATURES] for i in range(12)]
        frames[5] = pd.DataFrame([{**DEFAULTS, "credit_limit": 1e39}])[ALL_FEATURES]
        return await asyncio.gather(*[mb.score(f) for f in frames], return_exceptions=True)

    try:
        outs = asyncio.run(main())
    finally:
        mb.close()
    assert isinstance(outs[5], ValueError)
    for i, o in enumerate(outs):
        if i != 5:
            assert not isinstance(o, Exception) and np.allclose(o[0], [0.007])
    assert max(m.calls) > 1, "the requests were batched together first"


def test_422_body_of_a_non_object_row_equals_the_reference_model():
    from fastapi.exceptions import RequestValidationError
    from pydantic import TypeAdapter, ValidationError

    from databricks_kubernetes_mlops_poc_b200.ingest import parse_rows
    from databricks_kubernetes_mlops_poc_b200.schema import LoanApplicant

    ref = TypeAdapter(list[LoanApplicant])
    for raw in (b"[1]", b'[{"sex": 3}]', b'{"a": 1}', b"[[1]]", b'["x"]', b'[{"age": "old"}, 2]'):
        try:
            ref.validate_json(raw)
            want = None
        except ValidationError as e:
            want = [{**err, "loc": ("body", *err["loc"])} for err in e.errors(include_url=False, include_context=False)]
        try:
            parse_rows(raw)
            got = None
        except RequestValidationError as e:
            got = e.errors()
        assert got == want, raw


def test_log_pool_backlog_is_bounded():
    import threading

    from databricks_kubernetes_mlops_poc_b200.server import _BoundedLogPool

    pool = _BoundedLogPool(backlog=4)
    gate, done = threading.Event(), []
    for i in range(4):
        pool.submit(lambda i=i: (gate.wait(5), done.append(i)))
    assert pool.inline == 0
    pool.submit(done.append, "inline")  # backlog full: written on the caller's thread, nothing dropped
    assert pool.inline == 1 and done == ["inline"]
    gate.set()
    pool._pool.shutdown(wait=True)
    assert sorted(map(str, done)) == ["0", "1", "2", "3", "inline"]


# ----------------------------------------------------------------------------- row formats
def _pipeline_with_features(curated, cat, num):
    from sklearn.compose import ColumnTransformer
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.impute import SimpleImputer
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import OneHotEncoder

    catp = Pipeline([("imputer", SimpleImputer(strategy="constant", fill_value="missing")), ("ohe", OneHotEncoder(handle_unknown="ignore"))])
    nump = Pipeline([("imputer", SimpleImputer(strategy="median"))])
    pre = ColumnTransformer([("categorical", catp, cat), ("numeric", nump, num)])
    pipe = Pipeline([("preprocessor", pre), ("classifier", RandomForestClassifier(n_estimators=12, max_depth=5, random_state=0, n_jobs=-1))])
    tr = curated.iloc[:3000]
    pipe.fit(tr[cat + num], tr["default_payment_next_month"].to_numpy())
    return pipe


def test_packed_rows_need_exactly_nine_categoricals(curated):
    """The 64-byte packed row is decoded by the kernels as nine 7-bit fields + numerics from word 2: a schema with fewer
    categoricals must not be offered that layout (it takes 96-byte or ranked rows), and still walks correctly."""
    from blob_walk import walk_blob
    from rank_walk import walk_rank_layout

    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from oracle import reference_pipeline as rp

    cat, num = rp.CATEGORICAL_FEATURES[:7], rp.NUMERIC_FEATURES[:11]
    pipe = _pipeline_with_features(curated, cat, num)
    flat = flatten.flatten_pipeline(pipe)
    enc = RowEncoder(flat)
    assert not enc.packed_ok
    with pytest.raises(ValueError):
        enc.pack_rows(np.zeros((2, 24), dtype=np.uint32))
    df = curated[cat + num].iloc[3000:3700]
    want = pipe.predict_proba(df)[:, 1]
    rows = enc.encode_frame(df)  # > 128 rows: the native encoder, 96-byte rows
    assert np.abs(walk_blob(flat.blob, rows)[0] - want).max() <= 1e-12
    info = enc.rank_info()
    assert info.ok and info.row_bytes == 32  # 7 fields in 4 bytes + 11 uint16 -> 26 -> padded to 32
    p, _ = walk_rank_layout(enc.rank_layout(), info, flat.blob, enc.encode_frame_ranked(df))
    assert np.abs(p - want).max() <= 1e-12
    # an odd number of numerics: the one-hot values start at the next even pseudo-feature
    pipe2 = _pipeline_with_features(curated, cat, num[:9])
    flat2 = flatten.flatten_pipeline(pipe2)
    enc2 = RowEncoder(flat2)
    df2 = curated[cat + num[:9]].iloc[3000:3400]
    p2, _ = walk_rank_layout(enc2.rank_layout(), enc2.rank_info(), flat2.blob, enc2.encode_frame_ranked(df2))
    assert np.abs(p2 - pipe2.predict_proba(df2)[:, 1]).max() <= 1e-12


def test_frame_columns_reads_the_block_manager(curated, rf100d6):
    from databricks_kubernetes_mlops_poc_b200 import flatten
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES

    enc = RowEncoder(flatten.flatten_pipeline(rf100d6))
    df = curated[ALL_FEATURES].iloc[:1000]
    cols = enc.frame_columns(df)
    assert cols is not None
    scol, ptrs, strides, keep = cols
    x = df["credit_limit"].to_numpy()
    got = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_double * 1).from_address(ptrs[0]), shape=(1,))[0]
    assert got == x[0]
    # other column order, extra columns, an integer numeric column, a missing column, object strings
    df2 = df[ALL_FEATURES[::-1]].assign(extra=1)
    df2["age"] = df2["age"].astype(np.int64)
    assert enc.frame_columns(df2) is not None
    assert np.array_equal(enc.encode_frame_ranked(df2), enc.encode_frame_ranked(df))
    with pytest.raises(KeyError):
        enc.frame_columns(df.drop(columns=["age"]))
    obj = df.copy()
    obj["sex"] = obj["sex"].astype(object)
    assert enc.frame_columns(obj) is None  # None / NaN are distinct in object columns: the portable path decides


# ----------------------------------------------------------------------------- response lists
def test_recycled_float_lists_are_safe():
    from databricks_kubernetes_mlops_poc_b200 import _pylists as pl

    if not pl.available():
        pytest.skip("libb2fpy.so not built (no Python.h)")
    rng = np.random.default_rng(0)

    def build(a):
        b = pl.ListBuilder(len(a))
        for lo in range(0, len(a), 1000):
            b.fill(lo, a[lo:lo + 1000])
        return b.items

    a = rng.random(5000)
    held = build(a)
    snapshot = list(held)
    one = held[123]
    for _ in range(8):  # later responses while `held` (and one of its floats) is still referenced
        other = build(rng.random(5000))
        assert held == snapshot and one == snapshot[123]
    r0, f0 = pl.pool_stats()
    del other
    gc.collect()
    again = build(a * 3.0)
    assert again == (a * 3.0).tolist() and held == snapshot
    r1, f1 = pl.pool_stats()
    assert r1 > r0, "floats of dropped responses are recycled"
    rec = np.zeros(700, dtype=[("p", np.float64), ("l", np.int32), ("o", np.int32), ("s", np.float32), ("r", np.int32)])
    rec["p"], rec["o"] = rng.random(700), rng.integers(0, 2, 700)
    bp, bo = pl.ListBuilder(700), pl.ListBuilder(700)
    bp.fill(0, rec["p"])
    bo.fill(0, rec["o"])
    assert bp.items == rec["p"].tolist() and bo.items == rec["o"].tolist() and all(type(v) is int for v in bo.items)
    with pytest.raises(ValueError):
        pl.ListBuilder(10).fill(5, np.zeros(10))


def test_packed_block_encoder_every_block_shape(curated, rf100d6):
    """The 64-byte row encoder works on 256-row blocks, sixteen rows per transposing step (host_simd.cpp): every remainder
    shape, unaligned destinations, NaN / -0.0 / float32-overflow inputs and strided float64 columns agree with the portable path."""
    from databricks_kubernetes_mlops_poc_b200 import flatten, training
    from databricks_kubernetes_mlops_poc_b200.encode import RowEncoder
    from databricks_kubernetes_mlops_poc_b200.schema import ALL_FEATURES, NUMERIC_FEATURES

    flat = flatten.flatten_pipeline(rf100d6)
    enc, ref = RowEncoder(flat), RowEncoder(flat)
    ref._native_failed = True  # portable path only
    big = training.synth_frame(curated, 1200, seed=5, unknown_frac=0.05, nan_frac=0.1)[ALL_FEATURES]
    big.loc[big.index[3], "bill_amount_1"] = -0.0
    big.loc[big.index[700], "age"] = 3.0e38  # finite in float32
    for n in (1, 15, 16, 17, 255, 256, 257, 511, 1200):
        df = big.iloc[:n].reset_index(drop=True)
        want = ref.pack_rows(ref.encode_frame(df))
        for shift in (0, 1):  # 64-byte aligned destination (non-temporal stores) and an unaligned one
            buf = np.zeros(n * 16 + 32, dtype=np.uint32)
            off = (-buf.ctypes.data // 4) % 16 + shift
            got = buf[off:off + n * 16].reshape(n, 16)
            assert enc._encode_native(df, got, fmt=1)
            assert (got == want).all(), (n, shift)
    # float64 columns that are views into one 2-D block (element stride 14): the strided scalar conversion
    mat = np.ascontiguousarray(big[NUMERIC_FEATURES].to_numpy(dtype=np.float64))
    strided = big.copy()
    for k, name in enumerate(NUMERIC_FEATURES):
        strided[name] = mat[:, k]
    got = np.zeros((len(big), 16), dtype=np.uint32)
    assert enc._encode_native(strided, got, fmt=1)
    assert (got == ref.pack_rows(ref.encode_frame(big))).all()
    bad = big.copy()
    bad.loc[bad.index[1100], "payment_amount_3"] = 1e39  # overflows float32 -> sklearn's ValueError
    with pytest.raises(ValueError, match="infinity or a value too large"):
        enc._encode_native(bad, np.zeros((len(bad), 16), dtype=np.uint32), fmt=1)
    bad.loc[bad.index[1100], "payment_amount_3"] = -np.inf
    with pytest.raises(ValueError, match="infinity or a value too large"):
        enc._encode_native(bad, np.zeros((len(bad), 16), dtype=np.uint32), fmt=1)


def test_vector_vocabulary_lookup_matches_a_dictionary():
    """The eight-strings-per-step vocabulary lookup (host_simd.cpp: b2f_simd_hash_codes) against a Python dict: every
    string length around the 8- and 16-byte boundaries, unknown strings that share prefix, suffix and length with a
    vocabulary entry (only the middle differs), empty strings, the last strings of the buffer, int32 and int64 Arrow
    offsets, sliced arrays, and a column with nulls (scalar path)."""
    import ctypes as C

    import pyarrow as pa

    from databricks_kubernetes_mlops_poc_b200 import _cabi

    lib = _cabi.load_library()
    rng = np.random.default_rng(3)
    alphabet = np.array(list("abcdefghijklmnopqrstuvwxyz_0123456789"))

    def word(n):
        return "".join(rng.choice(alphabet, n))

    vocab = sorted({word(n) for n in (0, 1, 2, 5, 7, 8, 9, 12, 15, 16, 17, 18, 24, 33, 40) for _ in range(3)})
    longs = [w for w in vocab if len(w) > 16]
    twins = [w[:8] + word(len(w) - 16) + w[-8:] for w in longs]  # same (prefix, suffix, length), another middle
    twins = [t for t in twins if t not in vocab]
    unknown = [word(n) for n in (1, 3, 8, 9, 16, 17, 30)] + twins
    pool = np.array(vocab + [u for u in unknown if u not in vocab], dtype=object)
    want_of = {w: k for k, w in enumerate(vocab)}
    blob = "".join(vocab).encode()
    offs = np.cumsum([0] + [len(w.encode()) for w in vocab]).astype(np.int64)
    counts = np.array([len(vocab)], dtype=np.int32)
    null_codes = np.array([want_of[vocab[3]]], dtype=np.int32)
    enc = lib.b2f_encoder_create(1, 0, _cabi.ptr(counts), blob, _cabi.ptr(offs), _cabi.ptr(null_codes))
    assert enc
    try:
        for n in (1, 7, 8, 9, 64, 1000, 4099):
            values = pool[rng.integers(0, len(pool), n)]
            values[-1] = vocab[1] if len(vocab[1]) < 8 else values[-1]  # a short string at the very end of the buffer
            for typ in (pa.string(), pa.large_string()):
                for with_nulls in (False, True):
                    vals = list(values)
                    if with_nulls:
                        for k in range(0, n, 5):
                            vals[k] = None
                    arr = pa.array(vals, type=typ)
                    for sl in ((0, n), (min(3, n - 1), n)):
                        a = arr.slice(sl[0], sl[1] - sl[0])
                        validity, offsets, data = a.buffers()
                        col = (_cabi.StrColumn * 1)()
                        col[0].offsets = offsets.address
                        col[0].data = data.address if data is not None else 0
                        col[0].validity = validity.address if (validity is not None and a.null_count) else 0
                        col[0].offset = a.offset
                        col[0].data_bytes = data.size if data is not None else 0
                        col[0].offsets_are_64 = 1 if typ == pa.large_string() else 0
                        got = np.full(len(a), -99, dtype=np.int32)
                        assert lib.b2f_encoder_codes(enc, len(a), col, _cabi.ptr(got), 1) == 0
                        want = [int(null_codes[0]) if v is None else want_of.get(v, -1) for v in vals[sl[0]:sl[1]]]
                        assert got.tolist() == want, (n, str(typ), with_nulls, sl)
    finally:
        lib.b2f_encoder_destroy(enc)


def test_host_cpu_limit_and_default_pool_size():
    """The scorer's pool is sized from the container's CPU bandwidth (cgroup cpu.max): the C reading agrees with the bench's
    Python reading, and the default never exceeds quota - 2 (nor 48)."""
    import ctypes as C

    import bench
    from databricks_kubernetes_mlops_poc_b200 import _cabi

    lib = _cabi.load_library()
    limit = float(lib.b2f_host_cpu_limit())
    assert limit >= 0.0 and abs(limit - bench.cpu_bandwidth()) < 1e-9
    threads = lib.b2f_host_threads_default(0)  # without a GPU: the machine's CPUs instead of the GPU's NUMA node
    assert 1 <= threads <= 48
    if limit >= 3.0:
        assert threads <= int(limit) - 2
    ncpu = C.c_int(-1)
    node = lib.b2f_device_numa_node(0, C.byref(ncpu))  # no GPU here: -1 and 0 CPUs; on a GPU box: the node and its CPUs
    assert node >= -1 and ncpu.value >= 0
    assert bench.reference_procs("rf") == 1 and 1 <= bench.reference_procs("gbdt") <= 64
