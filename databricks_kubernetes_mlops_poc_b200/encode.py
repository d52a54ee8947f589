"""Host-side row encoder: request columns -> 96-byte encoded rows (layout in ``include/b2f.h``).

This is the host half of the "fused preprocess": everything sklearn's ColumnTransformer does
that needs *strings* happens here, vectorised per column, straight into the (pinned) staging
buffer; everything arithmetic (median imputation, one-hot comparison, float32 compare) happens
in the kernel.

Reference behaviour being matched (``databricks/src/01-train-model.ipynb:195-221``):

* ``SimpleImputer(constant "missing")`` + ``OneHotEncoder(handle_unknown="ignore")``: a category
  string is looked up in the sorted training vocabulary; unknown strings and missing values get
  code -1 (== all-zero one-hot block) unless "missing" itself was a training category;
* numeric columns are cast float64 -> float32 exactly as sklearn's tree predict does; NaN stays
  NaN (the kernel imputes the median); +-inf or a value that overflows float32 raises the same
  ``ValueError`` sklearn raises;
* columns are selected by NAME (``df[self.all_features]``, ``02-register-model.ipynb:335``), so any
  column order works (the reference's ``inference.csv`` puts ``credit_limit`` first).
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pandas as pd

try:  # Arrow's C++ hash lookup for the categorical columns; pandas fallback if absent
    import pyarrow as pa
    import pyarrow.compute as pc
except ImportError:  # pragma: no cover
    pa = pc = None

from .flatten import ROW_WORDS, FlatForest

PACKED_ROW_WORDS = 16  # B2F_ROWS_PACKED64: 64-byte rows (include/b2f.h)
NATIVE_THREADS = max(1, min(16, (os.cpu_count() or 1)))

_F32_MAX = float(np.finfo(np.float32).max)


class RowEncoder:
    def __init__(self, flat: FlatForest):
        self.cat_features = list(flat.cat_features)
        self.num_features = list(flat.num_features)
        self.n_cat = len(self.cat_features)
        self.n_num = len(self.num_features)
        self._index = [pd.Index(list(v), dtype=object) for v in flat.categories]
        self._lut = [{c: i for i, c in enumerate(v)} for v in flat.categories]
        self._colpos = {}  # column-order tuple -> positions of the model's features
        self._pa_vocab = [pa.array(list(v), type=pa.string()) for v in flat.categories] if pa is not None else None
        self._missing = list(flat.missing_codes) if flat.missing_codes else [-1] * self.n_cat  # NaN -> imputer constant
        self._none = list(flat.none_codes) if flat.none_codes else [-1] * self.n_cat  # None -> None category, if any
        # packed 64-byte rows: nine 7-bit (code + 1) fields + 14 float32 numerics
        self.packed_ok = self.n_cat <= 9 and self.n_num <= 14 and all(len(v) <= 126 for v in flat.categories)
        self._native = None  # b2f_encoder*, created on first use (needs libb200forest.so, not a GPU)
        self._native_failed = pa is None
        self._categories = [list(v) for v in flat.categories]

    # ------------------------------------------------------------------ native encoder (csrc/row_encoder.h)
    def _native_handle(self):
        if self._native is None and not self._native_failed:
            try:
                from . import _cabi

                lib = _cabi.load_library()
                blobs = [c.encode("utf-8") for v in self._categories for c in v]
                offsets = np.zeros(len(blobs) + 1, dtype=np.int64)
                np.cumsum([len(b) for b in blobs], out=offsets[1:])
                counts = np.array([len(v) for v in self._categories], dtype=np.int32)
                # a null entry: pandas' Arrow-backed strings cannot tell None from NaN; NaN semantics first
                nulls = np.array([m if m >= 0 else n for m, n in zip(self._missing, self._none)], dtype=np.int32)
                h = lib.b2f_encoder_create(self.n_cat, self.n_num, _cabi.ptr(counts), b"".join(blobs), _cabi.ptr(offsets), _cabi.ptr(nulls))
                if not h:
                    raise RuntimeError("b2f_encoder_create failed")
                self._native, self._lib, self._cabi = h, lib, _cabi
            except Exception:
                self._native_failed = True
        return self._native

    def __del__(self):
        try:
            if self._native:
                self._lib.b2f_encoder_destroy(self._native)
        except Exception:
            pass

    def _encode_native(self, df: pd.DataFrame, out: np.ndarray, packed: bool) -> bool:
        """Columnar fast path: Arrow string buffers + float64 columns -> rows, in C++ threads.  Returns False when
        a column does not have the expected physical type (the caller then takes the portable path)."""
        h = self._native_handle()
        if h is None:
            return False
        n = len(df)
        cols = (self._cabi.StrColumn * max(self.n_cat, 1))()
        keep = []  # keep the Arrow arrays alive during the call
        for j, name in enumerate(self.cat_features):
            ser = df[name]
            if ser.dtype == object:
                # object columns can hold None and NaN side by side, which the library treats differently
                # (None is not imputed); Arrow would merge them into one null -> portable path
                return False
            try:
                arr = pa.array(ser, from_pandas=True)
            except (pa.ArrowInvalid, pa.ArrowTypeError, pa.ArrowNotImplementedError):
                return False
            if isinstance(arr, pa.ChunkedArray):
                arr = arr.combine_chunks()
            if pa.types.is_dictionary(arr.type):
                arr = arr.dictionary_decode()
            if not (pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type)):
                return False
            validity, offsets, data = arr.buffers()
            keep.append((arr, validity, offsets, data))
            cols[j].offsets = offsets.address
            cols[j].data = data.address if data is not None else 0
            cols[j].validity = validity.address if (validity is not None and arr.null_count) else 0
            cols[j].offset = arr.offset
            cols[j].data_bytes = data.size if data is not None else 0
            cols[j].offsets_are_64 = 1 if pa.types.is_large_string(arr.type) else 0
        ptrs = (C.c_void_p * max(self.n_num, 1))()
        strides = np.ones(max(self.n_num, 1), dtype=np.int64)
        for k, name in enumerate(self.num_features):
            col = df[name].to_numpy()
            if col.dtype != np.float64:
                if col.dtype.kind not in "iuf":
                    return False
                col = col.astype(np.float64)
            keep.append(col)
            ptrs[k] = col.ctypes.data
            strides[k] = col.strides[0] // 8 if n > 1 else 1
        rc = self._lib.b2f_encoder_encode(h, n, cols, ptrs, self._cabi.ptr(strides), 1 if packed else 0, self._cabi.ptr(out), NATIVE_THREADS)
        if rc == -7:
            raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
        if rc != 0:
            raise RuntimeError(f"b2f_encoder_encode failed (rc={rc})")
        return True

    # ------------------------------------------------------------------ columns
    def encode_categorical(self, j: int, values) -> np.ndarray:
        """One categorical column (any array-like of str / None) -> int32 codes, -1 = unknown.
        Arrow's ``index_in`` does the vocabulary lookup in C++ (zero-copy for pandas' Arrow-backed
        string columns); non-string columns fall back to a pandas hash lookup."""
        if pa is not None:
            try:
                arr = pa.array(values, from_pandas=True)
                if pa.types.is_dictionary(arr.type):
                    arr = arr.dictionary_decode()
                if pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type):
                    idx = pc.index_in(arr, value_set=self._pa_vocab[j].cast(arr.type))
                    codes = idx.fill_null(-1).to_numpy(zero_copy_only=False).astype(np.int32)
                    if arr.null_count and (self._missing[j] >= 0 or self._none[j] >= 0):
                        codes[arr.is_null().to_numpy(zero_copy_only=False)] = self._null_codes(j, values)
                    return codes
            except (pa.ArrowInvalid, pa.ArrowTypeError, pa.ArrowNotImplementedError):
                pass
        arr = values if isinstance(values, (pd.Series, np.ndarray)) else np.asarray(values, dtype=object)
        codes = self._index[j].get_indexer(pd.Index(arr, dtype=object)).astype(np.int32)  # -1 = not in vocabulary
        if self._missing[j] >= 0 or self._none[j] >= 0:
            isna = np.asarray(pd.isna(arr))
            if isna.any():
                codes[isna] = self._null_codes(j, arr)
        return codes

    def _null_codes(self, j: int, values) -> np.ndarray:
        """Codes for the null entries of a column: None -> the None category (if fit saw one), NaN -> the
        imputer's constant category (if fit saw missing values); Arrow-backed columns only have one kind of null."""
        obj = np.asarray(values, dtype=object)
        nulls = obj[np.asarray(pd.isna(obj))]
        is_none = np.fromiter((v is None for v in nulls), dtype=bool, count=len(nulls))
        return np.where(is_none, self._none[j], self._missing[j]).astype(np.int32)

    @staticmethod
    def cast_numeric(block64: np.ndarray) -> np.ndarray:
        """float64 (N, k) -> float32 with sklearn's finiteness rule (NaN allowed: missing)."""
        bad = np.abs(block64) > _F32_MAX  # False for NaN
        if bad.any():
            raise ValueError("Input X contains infinity or a value too large for dtype('float32').")
        return block64.astype(np.float32)

    # ------------------------------------------------------------------ frames
    SMALL_BATCH = 128  # below this, per-column vectorised machinery costs more than a Python loop

    def encode_frame(self, df: pd.DataFrame, out: np.ndarray | None = None) -> np.ndarray:
        """DataFrame with (at least) the 23 named columns -> uint32 (N, 24) encoded rows."""
        n = len(df)
        if out is None:
            out = np.empty((n, ROW_WORDS), dtype=np.uint32)
        else:
            assert out.shape == (n, ROW_WORDS) and out.dtype == np.uint32
        cols = df.columns
        missing = [c for c in self.cat_features + self.num_features if c not in cols]
        if missing:
            raise KeyError(f"{missing} not in index")  # what df[self.all_features] raises
        as_i32 = out.view(np.int32)
        if n <= self.SMALL_BATCH:
            # small request: ONE object-array extraction of the whole frame (column selection alone costs
            # pandas ~0.5 ms), then dictionary lookups / float() in Python
            key = tuple(cols)
            pos = self._colpos.get(key)
            if pos is None:
                pos = self._colpos[key] = [cols.get_loc(c) for c in self.cat_features + self.num_features]
            recs = df.to_numpy(dtype=object)
            nums = np.empty((n, self.n_num), dtype=np.float64)
            for i in range(n):
                rec = recs[i]
                for j in range(self.n_cat):
                    v = rec[pos[j]]
                    as_i32[i, j] = self._lut[j].get(v, -1) if isinstance(v, str) else (self._none[j] if v is None else (self._missing[j] if v != v else -1))
                for k in range(self.n_num):
                    v = rec[pos[self.n_cat + k]]
                    nums[i, k] = np.nan if v is None else float(v)  # float("abc") raises ValueError, as pd.to_numeric does
            if self.n_num:
                out.view(np.float32)[:, self.n_cat : self.n_cat + self.n_num] = self.cast_numeric(nums)
        elif self._encode_native(df, out, packed=False):
            return out
        else:
            for j, name in enumerate(self.cat_features):
                as_i32[:, j] = self.encode_categorical(j, df[name])
            if self.n_num:
                try:
                    block = df[self.num_features].to_numpy(dtype=np.float64, na_value=np.nan)
                except (ValueError, TypeError):
                    block = np.empty((n, self.n_num), dtype=np.float64)
                    for k, name in enumerate(self.num_features):
                        block[:, k] = pd.to_numeric(df[name], errors="raise").to_numpy(dtype=np.float64, na_value=np.nan)
                out.view(np.float32)[:, self.n_cat : self.n_cat + self.n_num] = self.cast_numeric(block)
        out[:, self.n_cat + self.n_num :] = 0
        return out

    def encode_arrays(self, codes: np.ndarray, nums: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """Already-dictionary-encoded input (int codes (N, n_cat), float nums (N, n_num)) -> rows."""
        n = codes.shape[0]
        if out is None:
            out = np.empty((n, ROW_WORDS), dtype=np.uint32)
        out.view(np.int32)[:, : self.n_cat] = codes
        out.view(np.float32)[:, self.n_cat : self.n_cat + self.n_num] = self.cast_numeric(
            np.asarray(nums, dtype=np.float64)
        )
        out[:, self.n_cat + self.n_num :] = 0
        return out

    # ------------------------------------------------------------------ packed 64-byte rows
    def pack_rows(self, rows24: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """(N, 24) encoded rows -> (N, 16) packed rows (B2F_ROWS_PACKED64): one third fewer bytes over PCIe.
        Lossless: code + 1 in 7 bits (0 = unknown), numerics untouched."""
        if not self.packed_ok:
            raise ValueError("schema does not fit the packed row (<= 9 categoricals of <= 126 categories, <= 14 numerics)")
        n = rows24.shape[0]
        if out is None:
            out = np.zeros((n, PACKED_ROW_WORDS), dtype=np.uint32)
        fields = (rows24.view(np.int32)[:, : self.n_cat].astype(np.int64) + 1).astype(np.uint64)
        word = np.zeros(n, dtype=np.uint64)
        for j in range(self.n_cat):
            word |= fields[:, j] << np.uint64(7 * j)
        out[:, 0] = (word & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        out[:, 1] = (word >> np.uint64(32)).astype(np.uint32)
        out[:, 2 : 2 + self.n_num] = rows24[:, self.n_cat : self.n_cat + self.n_num]
        out[:, 2 + self.n_num :] = 0
        return out

    def encode_frame_packed(self, df: pd.DataFrame, out: np.ndarray | None = None) -> np.ndarray:
        """DataFrame -> (N, 16) packed rows; large frames go through the native encoder in one pass."""
        n = len(df)
        if n > self.SMALL_BATCH and self.packed_ok:
            missing = [c for c in self.cat_features + self.num_features if c not in df.columns]
            if missing:
                raise KeyError(f"{missing} not in index")
            if out is None:
                out = np.empty((n, PACKED_ROW_WORDS), dtype=np.uint32)
            if self._encode_native(df, out, packed=True):
                return out
        return self.pack_rows(self.encode_frame(df), out=out)

    def encode_arrays_packed(self, codes: np.ndarray, nums: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        return self.pack_rows(self.encode_arrays(codes, nums), out=out)
