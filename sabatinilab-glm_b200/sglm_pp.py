"""
sglm_pp — drop-in for the reference module `backend/sglm_pp.py`.

Hot path (GPU, libsglm_b200.so): `timeshift`, `timeshift_multiple`, `shift`,
`concat_start_crop_end`, `concat_end_crop_start` — the lag/shift design-matrix builder
(reference backend/sglm_pp.py:23-103, :298-357, :436-486).  The whole T x (P*L) design is
produced by ONE gather launch from a column map (source column, shift) instead of >= 4
full-size host copies per shift block; the result is bit-exact, NaN/edge padding included.

Inputs: numpy ndarray or pandas DataFrame (type-preserving, as in the reference), or a
CUDA torch tensor (then a CUDA tensor is returned and nothing leaves the device).

The small pandas-typed helpers (`zscore`, `diff`, `get_column_nums`, bucket / CV index
builders, `detrend_data`) are not kernels (CPU passthrough, out of scope);
they keep the reference's semantics, including `bucket_ids_by_timeframe` dividing by the
number of buckets (backend/sglm_pp.py:232-233).
"""
from typing import List, Optional, Union

import numpy as np
import pandas as pd

import _engine as eng
import _sglm_native as nat


# --------------------------------------------------------------------------- #
# column-map construction (host logic)
# --------------------------------------------------------------------------- #
def _resolve_inx(n_cols, shift_inx):
    inx = list(range(n_cols)) if len(shift_inx) == 0 else [int(i) for i in shift_inx]
    out = []
    for i in inx:
        if i < -n_cols or i >= n_cols:
            raise IndexError(f"index {i} is out of bounds for axis 1 with size {n_cols}")
        out.append(i % n_cols)
    return out


def build_column_map(n_cols, shift_inx, shift_amt_list, unshifted_keep_all=True):
    """(col_src, col_shift, block_sizes): shift-major, predictor-minor; the zero-shift block
    holds ALL columns when `unshifted_keep_all` (backend/sglm_pp.py:83-89, :436-457)."""
    inx = _resolve_inx(n_cols, shift_inx)
    src, sh, sizes = [], [], []
    for a in shift_amt_list:
        a = int(a)
        if a == 0 and unshifted_keep_all:
            src.extend(range(n_cols)); sh.extend([0] * n_cols); sizes.append(n_cols)
        else:
            src.extend(inx); sh.extend([a] * len(inx)); sizes.append(len(inx))
    return np.asarray(src, dtype=np.int32), np.asarray(sh, dtype=np.int32), sizes


def _to_f64_source(vals):
    """Host array -> float64 words for the gather.  8-byte integer types are moved as raw
    bits when no fill can occur (all shifts zero); everything else is converted (exact for
    the value ranges the reference's float64 result can hold)."""
    return np.ascontiguousarray(vals, dtype=np.float64)


def _gather_host(vals, col_src, col_shift, fill_value):
    """numpy in -> numpy out through the GPU gather; dtype rule of the reference: any
    non-zero shift makes the block float64 (blanks are float64, :312), pure zero-shift
    selections keep the input dtype."""
    all_zero = not np.any(col_shift)
    if all_zero and vals.dtype != np.float64:
        if vals.dtype.itemsize == 8 and vals.dtype.kind in "iu":
            words = np.ascontiguousarray(vals).view(np.float64)       # opaque 8-byte words
            out = eng.gather(eng.device_matrix(words), col_src, col_shift, 0.0).cpu().numpy()
            return out.view(vals.dtype)
        out = eng.gather(eng.device_matrix(_to_f64_source(vals)), col_src, col_shift, 0.0).cpu().numpy()
        return out.astype(vals.dtype)
    Xd = eng.device_matrix(_to_f64_source(vals))
    return eng.gather(Xd, col_src, col_shift, fill_value).cpu().numpy()


def _restore_dtypes(res, X, src, sh):
    """Un-shifted columns are copies of X's columns and keep their dtype (X.copy() in
    backend/sglm_pp.py:403); shifted columns are float64."""
    for i in np.flatnonzero(np.asarray(sh) == 0):
        dt = X.dtypes.iloc[int(src[i])]
        if res.dtypes.iloc[int(i)] != dt:
            res.isetitem(int(i), res.iloc[:, int(i)].astype(dt))
    return res


# --------------------------------------------------------------------------- #
# device-resident design (numpy / pandas callers without the PCIe round trip)
# --------------------------------------------------------------------------- #
def _device_default():
    import os
    return os.environ.get("SGLM_DEVICE_DESIGN", "0") == "1"


class DeviceDesign:
    """Lag design that lives on the GPU: what `timeshift_multiple(..., device=True)` returns for
    numpy / pandas input (or always with SGLM_DEVICE_DESIGN=1).

    The reference hands the T x (P*L) design back as a host array, the drivers `dropna()` it, pick the
    predictor and response columns and pass `.values` to the fit (er_refactored_from_scratch_cleanup.py:
    421-452; backend/test/test_sglm_ez.py:23-41).  At 2M x 2000 that is a 32 GB matrix crossing PCIe
    twice.  This object keeps the *recipe* instead — base signals on the device + column map + row
    selection — and supports exactly those steps lazily:

        d = sglm_pp.timeshift_multiple(X, ..., device=True)     # or sglm_ez.timeshift_cols(df, ..., device=True)
        d = d.dropna()               # valid rows decided from the base signals (sglm_lag_valid_rows)
        Xd, yd = d[x_cols], d[y_col] # column subsets of the map: free
        sglm_cv.cv_glm_mult_params(Xd, yd, cv_idx, ...)          # gather runs here, on the device, valid rows only

    Anything else sees an ordinary array: `np.asarray(d)`, `d.values`, `d.to_numpy()`, `d.to_pandas()`
    materialise on the host (bit-identical to the reference's result)."""

    def __init__(self, base, col_src, col_shift, fill_value, names=None, index=None, rows=None, n_base_rows=None):
        self._base = base                                   # CUDA float64 [T, P]
        self._src = np.ascontiguousarray(col_src, dtype=np.int32)
        self._sh = np.ascontiguousarray(col_shift, dtype=np.int32)
        self._fill = fill_value
        self._names = None if names is None else list(names)
        self._index = index                                 # pandas index of the base rows (or None)
        self._rows = rows                                   # None (all rows) | (lo, hi) | CUDA int64 row list
        self._T = int(base.shape[0]) if n_base_rows is None else int(n_base_rows)
        self._cache = None

    # ---- shape / labels
    @property
    def shape(self):
        if self._rows is None:
            n = self._T
        elif isinstance(self._rows, tuple):
            n = self._rows[1] - self._rows[0]
        else:
            n = int(self._rows.numel())
        return (n, int(self._src.shape[0]))

    def __len__(self):
        return self.shape[0]

    @property
    def ndim(self):
        return 2

    @property
    def dtype(self):
        return np.dtype(np.float64)

    @property
    def columns(self):
        return pd.Index(self._names) if self._names is not None else pd.RangeIndex(self.shape[1])

    def row_positions(self):
        """Positions (in the base signals) of the rows of this design, as a host int64 array."""
        if self._rows is None:
            return np.arange(self._T, dtype=np.int64)
        if isinstance(self._rows, tuple):
            return np.arange(self._rows[0], self._rows[1], dtype=np.int64)
        return self._rows.cpu().numpy()

    @property
    def index(self):
        pos = self.row_positions()
        return self._index[pos] if self._index is not None else pd.RangeIndex(self._T)[pos]

    # ---- lazy operations
    def _derive(self, src=None, sh=None, names=None, rows="same"):
        d = DeviceDesign(self._base, self._src if src is None else src, self._sh if sh is None else sh, self._fill,
                         self._names if names is None else names, self._index,
                         self._rows if isinstance(rows, str) else rows, self._T)
        return d

    def __getitem__(self, key):
        """Column selection by name(s) (or integer positions for unnamed designs)."""
        single = not isinstance(key, (list, tuple, np.ndarray, pd.Index))
        keys = [key] if single else list(key)
        if self._names is not None:
            pos = []
            for k in keys:
                hits = [i for i, n in enumerate(self._names) if n == k]
                if not hits:
                    raise KeyError(k)
                pos.extend(hits)
        else:
            pos = [int(k) for k in keys]
        names = [self._names[i] for i in pos] if self._names is not None else None
        out = self._derive(self._src[pos], self._sh[pos], names)
        if names is None:
            out._names = None
        return out

    def dropna(self):
        """Rows without NaN (pandas `dropna()` / `X[~np.isnan(X).any(axis=1)]`), decided on the device from
        the base signals and the column map — the dropped rows are never built."""
        torch = nat.require_cuda()
        T = self._T
        valid = torch.empty(max(T, 1), dtype=torch.uint8, device="cuda")
        summ = torch.empty(3, dtype=torch.int64, device="cuda")
        C = int(self._src.shape[0])
        both = eng._dev(np.concatenate([self._src, self._sh]), np.int32)
        fill_nan = int(isinstance(self._fill, float) and np.isnan(self._fill)) if not hasattr(self._fill, "dtype") else int(np.isnan(self._fill))
        nat.call("sglm_lag_valid_rows", nat.ptr(self._base), T, int(self._base.shape[1]), eng.row_stride(self._base),
                 nat.ptr(both[:C]), nat.ptr(both[C:]), C, int(self._sh.min()) if C else 0, int(self._sh.max()) if C else 0,
                 fill_nan, nat.ptr(valid), nat.ptr(summ), nat.stream_ptr())
        if self._rows is not None:                          # an earlier selection: intersect
            keep = torch.zeros(max(T, 1), dtype=torch.uint8, device="cuda")
            if isinstance(self._rows, tuple):
                keep[self._rows[0]:self._rows[1]] = 1
            else:
                keep[self._rows] = 1
            valid &= keep
            summ = torch.stack([valid[:T].sum(dtype=torch.int64), torch.tensor(T, device="cuda"), torch.tensor(-1, device="cuda")])
            nz = torch.nonzero(valid[:T]).reshape(-1)
            return self._derive(rows=nz)
        cnt, first, last = (int(v) for v in summ.cpu().numpy())
        if cnt == 0:
            return self._derive(rows=(0, 0))
        if last - first + 1 == cnt:
            return self._derive(rows=(first, last + 1))
        rows = torch.empty(cnt, dtype=torch.int64, device="cuda")
        wb = nat.lib().sglm_mask_compact_workspace_bytes(T)
        ws = torch.empty((wb + 7) // 8, dtype=torch.int64, device="cuda")
        nat.call("sglm_mask_compact_rows", nat.ptr(valid), T, nat.ptr(rows), nat.ptr(ws), ws.numel() * 8, nat.stream_ptr())
        return self._derive(rows=rows)

    def copy(self):
        return self._derive()

    def lag_recipe(self):
        """The design as an engine-level recipe (base signals + column map + contiguous row range) when its rows
        ARE a contiguous range — then the CV grid computes its statistics from the base signals and the design is
        never built; None otherwise (row lists after an interior NaN: the gathered tensor is used)."""
        if self._rows is not None and not isinstance(self._rows, tuple):
            return None
        lo, hi = (0, self._T) if self._rows is None else self._rows
        return eng.LagRecipe(self._base, self._src, self._sh, lo, hi, self._fill, cache=self._cache)

    # ---- materialisation
    def tensor(self):
        """The design as a CUDA float64 tensor [rows, C] (built once, cached): plain gather + row view when
        the rows are a contiguous range, row-list gather otherwise."""
        if self._cache is None:
            torch = nat.require_cuda()
            if self._rows is None or isinstance(self._rows, tuple):
                full = eng.gather(self._base, self._src, self._sh, self._fill)
                self._cache = full if self._rows is None else full[self._rows[0]:self._rows[1]]
            else:
                n, C = self.shape
                out = torch.empty((n, C), dtype=torch.float64, device="cuda")
                if n and C:
                    both = eng._dev(np.concatenate([self._src, self._sh]), np.int32)
                    nat.call("sglm_timeshift_rows_f64", nat.ptr(self._base), self._T, int(self._base.shape[1]),
                             eng.row_stride(self._base), nat.ptr(both[:C]), nat.ptr(both[C:]), C,
                             nat.f64_bits(self._fill), nat.ptr(self._rows), n, nat.ptr(out), C, nat.stream_ptr())
                self._cache = out
        return self._cache

    def to_numpy(self, dtype=None, copy=False):
        a = self.tensor().cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    def __array__(self, dtype=None, copy=None):
        return self.to_numpy(dtype)

    @property
    def values(self):
        return self.to_numpy()

    def to_pandas(self):
        return pd.DataFrame(self.to_numpy(), index=self.index, columns=self.columns)

    def __repr__(self):
        return f"DeviceDesign(shape={self.shape}, on cuda; np.asarray() / .to_pandas() materialise it)"


# --------------------------------------------------------------------------- #
# hot path
# --------------------------------------------------------------------------- #
def timeshift(X, shift_inx=[], shift_amt=1, keep_non_inx=False, dct=None, fill_value=np.nan):
    """Shift the columns `shift_inx` of X down (shift_amt > 0) or up (< 0); vacated rows get
    `fill_value`.  Reference: backend/sglm_pp.py:23-56."""
    is_df = type(X) == pd.DataFrame
    is_t = eng.is_torch(X)
    n_cols = X.shape[1]
    inx = _resolve_inx(n_cols, shift_inx)
    a = int(shift_amt)
    if keep_non_inx:
        src = np.arange(n_cols, dtype=np.int32)
        sh = np.zeros(n_cols, dtype=np.int32)
        sh[inx] = a
    else:
        src = np.asarray(inx, dtype=np.int32)
        sh = np.full(len(inx), a, dtype=np.int32)
    if is_t:
        res = eng.gather(eng.device_matrix(X), src, sh, fill_value)
    elif is_df:
        vals = _gather_host(X.values, src, sh, fill_value)
        cols = X.columns if keep_non_inx else X.columns[inx]
        res = _restore_dtypes(pd.DataFrame(vals, index=X.index, columns=cols), X, src, sh)
    else:
        vals = np.asarray(X)
        res = _gather_host(vals, src, sh, fill_value)
        if keep_non_inx and res.dtype != vals.dtype:
            with np.errstate(invalid="ignore"):
                res = res.astype(vals.dtype)      # X.copy()[:, inx] = shifted keeps X's dtype (:430-431)
    if dct is not None:
        dct[shift_amt] = res
    return res


def timeshift_multiple(X, shift_inx=[], shift_amt_list=[-1, 0, 1], unshifted_keep_all=True, fill_value=np.nan,
                       device=None):
    """All shifts of `shift_amt_list` as column blocks of one array, in list order; the
    zero-shift block keeps every column when `unshifted_keep_all`.  DataFrames get the
    reference's names: `col` for shift 0, f"{col}_{shift}" otherwise.
    Reference: backend/sglm_pp.py:58-103, :436-486 (one thread and >= 4 copies per shift)."""
    is_df = type(X) == pd.DataFrame
    n_cols = X.shape[1]
    shift_amt_list = [int(a) if float(a).is_integer() else a for a in shift_amt_list]
    src, sh, sizes = build_column_map(n_cols, shift_inx, shift_amt_list, unshifted_keep_all)
    if len(shift_amt_list) == 0:
        raise ValueError("need at least one array to concatenate")
    if eng.is_torch(X):
        if device:         # CUDA tensor in, lazy device-resident design out (explicit opt-in only)
            return DeviceDesign(eng.device_matrix(X), src, sh, fill_value, None, None)
        return eng.gather(eng.device_matrix(X), src, sh, fill_value)
    if device is None:
        device = _device_default()
    names = None
    if is_df:
        names, pos = [], 0
        for a, n in zip(shift_amt_list, sizes):
            base = X.columns[src[pos:pos + n]]
            names.extend(list(base) if a == 0 else [f"{c}_{a}" for c in base])
            pos += n
    if device:
        # extension: the design stays on the GPU as a lazy DeviceDesign (see its docstring)
        base_vals = X.values if is_df else np.asarray(X)
        return DeviceDesign(eng.device_matrix(_to_f64_source(base_vals)), src, sh, fill_value, names,
                            X.index if is_df else None)
    if is_df:
        vals = _gather_host(X.values, src, sh, fill_value)
        return _restore_dtypes(pd.DataFrame(vals, index=X.index, columns=names), X, src, sh)
    return _gather_host(np.asarray(X), src, sh, fill_value)


def shift(setup_array: np.ndarray, shift_amt: int, fill_value: Optional[float] = np.nan):
    """Shift every column of `setup_array` (backend/sglm_pp.py:298-319); a zero shift returns
    the input object itself, as the reference does."""
    if shift_amt == 0:
        return setup_array
    n = setup_array.shape[1]
    src = np.arange(n, dtype=np.int32)
    sh = np.full(n, int(shift_amt), dtype=np.int32)
    if eng.is_torch(setup_array):
        return eng.gather(eng.device_matrix(setup_array), src, sh, fill_value)
    return _gather_host(np.asarray(setup_array), src, sh, fill_value)


def _concat_crop(blanks, X_to_shift, at_start):
    torch = nat.require_cuda()
    is_t = eng.is_torch(X_to_shift)
    Xd = eng.device_matrix(X_to_shift)
    Bd = eng.device_matrix(blanks)
    T, n = Xd.shape
    k = Bd.shape[0]
    if Bd.shape[1] != n:
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    out = torch.empty((T, n), dtype=torch.float64, device="cuda")
    kk = min(k, T)

    def crop(srcT, row_begin, rows, dst_row):
        if rows > 0 and n > 0:
            dst = out[dst_row:]
            nat.call("sglm_crop_rows_f64", nat.ptr(srcT), eng.row_stride(srcT), row_begin, rows, n,
                     nat.ptr(dst), n, nat.stream_ptr())
    if k == 0:
        # reference quirk: `[:-0]` is an empty slice, `[0:]` is everything (:337, :356)
        res = out[:0] if at_start else Xd.clone()
        return res if is_t else res.cpu().numpy()
    if at_start:      # rows of concat([blanks, X])[:-k]  == first T rows
        crop(Bd, 0, kk, 0)
        crop(Xd, 0, T - kk, kk)
    else:             # rows of concat([X, blanks])[k:]   == last T rows
        crop(Xd, kk, T - kk, 0)
        crop(Bd, k - kk, kk, T - kk)
    return out if is_t else out.cpu().numpy()


def concat_start_crop_end(blanks: np.ndarray, X_to_shift: np.ndarray):
    """concat([blanks, X])[:-len(blanks)] (backend/sglm_pp.py:321-338) on the device."""
    return _concat_crop(blanks, X_to_shift, True)


def concat_end_crop_start(blanks: np.ndarray, X_to_shift: np.ndarray):
    """concat([X, blanks])[len(blanks):] (backend/sglm_pp.py:340-357) on the device."""
    return _concat_crop(blanks, X_to_shift, False)


# --------------------------------------------------------------------------- #
# type adaptors kept for API compatibility (backend/sglm_pp.py:280-296, :359-486)
# --------------------------------------------------------------------------- #
def get_numpy_version(X: Union[np.ndarray, pd.DataFrame]) -> np.ndarray:
    return X.values if type(X) == pd.DataFrame else X


def shifted_cols_to_pandas(X, shifted_X, shift_inx, keep_non_inx):
    shift_inx = list(shift_inx)
    out = X.copy()
    vals = np.asarray(shifted_X)
    for k, c in enumerate(shift_inx):
        out.isetitem(c, vals[:, k])
    return out if keep_non_inx else out.iloc[:, shift_inx]


def shifted_cols_to_numpy(X, shifted_X, shift_inx, keep_non_inx):
    if not keep_non_inx:
        return np.array(shifted_X, copy=True)
    out = X.copy()
    with np.errstate(invalid="ignore"):
        out[:, list(shift_inx)] = shifted_X
    return out


def shifted_cols_to_original_type(X, shifted_X, shift_inx, keep_non_inx):
    fn = shifted_cols_to_pandas if type(X) == pd.DataFrame else shifted_cols_to_numpy
    return fn(X, shifted_X, shift_inx, keep_non_inx)


def concat_pandas_shifts(shift_amt_list: List[int], shifted_list):
    frames = []
    for a, blk in zip(shift_amt_list, shifted_list):
        frames.append(blk if a == 0 else blk.rename({c: f"{c}_{a}" for c in blk.columns}, axis=1))
    return pd.concat(frames, axis=1)


def concat_all_shifts(X, shift_amt_list, shifted_list):
    if type(X) == pd.DataFrame:
        return concat_pandas_shifts(shift_amt_list, shifted_list)
    if len(shifted_list) and eng.is_torch(shifted_list[0]):
        import torch
        return torch.cat(list(shifted_list), dim=1)
    return np.concatenate(shifted_list, axis=1)


# --------------------------------------------------------------------------- #
# CPU passthrough helpers (not kernels; semantics of backend/sglm_pp.py:105-264, :488-545)
# --------------------------------------------------------------------------- #
def zscore(X):
    """(X - mean) / std along axis 0 (backend/sglm_pp.py:105-118).  numpy / pandas input: the reference's expression
    on the host (numpy std has ddof 0, pandas std ddof 1 and skips NaN).  CUDA tensor input (extension, SURVEY.md
    §8f-3): the same with numpy's convention on the device (sglm_col_moments_f64 + sglm_zscore_apply_f64)."""
    if eng.is_torch(X):
        return zscore_device(X, ddof=0, skipna=False)
    return (X - X.mean(axis=0)) / X.std(axis=0)


def zscore_device(X, ddof=0, skipna=False):
    """z-score the columns of a CUDA tensor; ddof / skipna select numpy (0, False) or pandas (1, True) conventions."""
    torch = nat.require_cuda()
    one_d = X.dim() == 1
    Xd = eng.device_matrix(X)
    T, C = Xd.shape
    mean = torch.empty(C, dtype=torch.float64, device="cuda")
    sd = torch.empty(C, dtype=torch.float64, device="cuda")
    wb = nat.lib().sglm_col_moments_workspace_bytes(T, C)
    ws = torch.empty(wb // 8 + 1, dtype=torch.float64, device="cuda")
    nat.call("sglm_col_moments_f64", nat.ptr(Xd), eng.row_stride(Xd), T, C, int(ddof), int(bool(skipna)), nat.ptr(mean),
             nat.ptr(sd), nat.ptr(ws), ws.numel() * 8, nat.stream_ptr())
    out = torch.empty((T, C), dtype=torch.float64, device="cuda")
    nat.call("sglm_zscore_apply_f64", nat.ptr(Xd), eng.row_stride(Xd), T, C, nat.ptr(mean), nat.ptr(sd), nat.ptr(out), C,
             nat.stream_ptr())
    return out[:, 0] if one_d else out


def diff_device(X, diff_inx=[], n=1, append_to_base=False, fill_value=np.nan):
    """n-th difference along axis 0 of the chosen columns of a CUDA tensor (np.diff semantics: n repeated first
    differences, bit-exact); append_to_base pads n fill rows on top and appends the columns to X as `diff` does."""
    torch = nat.require_cuda()
    Xd = eng.device_matrix(X)
    T, C = Xd.shape
    cols = [int(c) for c in diff_inx] if len(diff_inx) else list(range(C))
    cols_d = eng._dev(cols, np.int32)
    cur, cur_cols = Xd, cols_d
    for k in range(int(n)):
        rows = cur.shape[0]
        nxt = torch.empty((max(rows - 1, 0), len(cols)), dtype=torch.float64, device="cuda")
        if rows > 1:
            nat.call("sglm_diff1_f64", nat.ptr(cur), eng.row_stride(cur), rows, nat.ptr(cur_cols) if k == 0 else None,
                     len(cols), nat.ptr(nxt), len(cols), nat.stream_ptr())
        cur, cur_cols = nxt, None
    if not append_to_base:
        return cur
    pad = torch.full((int(n), len(cols)), float(fill_value), dtype=torch.float64, device="cuda")
    return torch.cat([Xd, torch.cat([pad, cur], dim=0)], dim=1)


def diff(X, diff_inx=[], n=1, axis=0, append_to_base=False, fill_value=np.nan, **kwargs):
    """n-th difference of the chosen columns (backend/sglm_pp.py:120-190).  CUDA tensors take the device kernel."""
    if eng.is_torch(X):
        return diff_device(X, diff_inx, n=n, append_to_base=append_to_base, fill_value=fill_value)
    out_type = type(X)
    if out_type == pd.Series and append_to_base:
        out_type = pd.DataFrame
    if type(X) == pd.Series:
        X = pd.DataFrame(X)
    cols = list(diff_inx) if diff_inx else list(range(X.shape[1]))
    framed = type(X) == pd.DataFrame
    if framed:
        names = [str(c) + '_diff' for c in X.columns[cols]]
        if append_to_base:
            names = list(X.columns) + names
        vals = X.values
    else:
        vals = X
    if len(X.shape) == 1:
        vals = vals.reshape((-1, 1))
    res = np.diff(vals[:, cols], n=n, axis=axis, **kwargs)
    if append_to_base:
        res = np.concatenate([np.full((n, res.shape[1]), fill_value, dtype=np.float64), res], axis=0)
        res = np.concatenate([vals, res], axis=-1)
        index = X.index if framed else None
    elif framed:
        index = X.index[1:]
    if framed:
        res = pd.DataFrame(res, columns=names, index=index)
    if out_type == pd.Series:
        res = res.iloc[:, 0]
    return res


def get_column_nums(df, column_names=[]):
    """Positions of the named columns; duplicate names are an error (backend/sglm_pp.py:192-209)."""
    locs = [df.columns.get_loc(name) for name in column_names]
    if any(type(loc) == np.ndarray for loc in locs):
        raise ValueError('Duplicate column found in X column names.')
    return locs


def bucket_ids_by_timeframe(total_timesteps, timesteps_per_bucket=20):
    """Bucket id per timestep.  Reproduces the reference's arithmetic, which divides by the
    NUMBER of buckets rather than by the bucket length (backend/sglm_pp.py:232-233)."""
    num_buckets = total_timesteps // timesteps_per_bucket
    return np.arange(total_timesteps) // num_buckets


def cv_idx_from_bucket_ids(bucket_ids, X, y=None, num_folds=None, test_size=None, device=None):
    """GroupShuffleSplit over the bucket ids (backend/sglm_pp.py:236-264); global numpy RNG
    state decides the split, as in the reference.

    device="cuda" (extension): the same splits as CUDA int64 index tensors.  The
    group-level shuffle — the only random part — runs on the host exactly as scikit-learn's
    GroupShuffleSplit does it (ShuffleSplit over the sorted unique groups, same draws from the global
    numpy RNG); the expansion to the (millions of) row indices happens on the device, so the index
    lists never cross the PCIe bus and `cv_glm_*` takes them as they are."""
    from sklearn.model_selection import GroupShuffleSplit
    if num_folds is None:
        num_folds = bucket_ids.max() + 1
    if test_size is None:
        test_size = 1 / num_folds
    if device is None:
        splitter = GroupShuffleSplit(n_splits=num_folds, test_size=test_size)
        return list(splitter.split(X, y, bucket_ids))
    import torch
    from sklearn.model_selection import ShuffleSplit
    ids = bucket_ids if type(bucket_ids).__module__.startswith("torch") else torch.from_numpy(np.array(bucket_ids, copy=True))
    ids = ids.to(device)
    classes, inverse = torch.unique(ids, sorted=True, return_inverse=True)
    n_classes = int(classes.numel())
    out = []
    # sklearn/model_selection/_split.py: GroupShuffleSplit._iter_indices = ShuffleSplit._iter_indices over the classes
    for g_train, g_test in ShuffleSplit(n_splits=int(num_folds), test_size=test_size).split(np.empty((n_classes, 1))):
        lut = torch.zeros(n_classes, dtype=torch.bool, device=ids.device)
        lut[torch.from_numpy(g_test).to(ids.device)] = True
        in_test = lut[inverse]
        lut.zero_()
        lut[torch.from_numpy(g_train).to(ids.device)] = True
        in_train = lut[inverse]
        out.append((torch.nonzero(in_train).reshape(-1), torch.nonzero(in_test).reshape(-1)))
    return out


def min_max_scale(X, lower_bound, upper_bound):
    return (X - lower_bound) / (upper_bound - lower_bound)


def lambda_min_max(X: pd.Series) -> float:
    centre = X.iloc[(len(X) + 1) // 2 - 1]
    return min_max_scale(centre, X.quantile(0.05), X.quantile(0.95))


def detrend_data(X: pd.DataFrame, detrend_col: str, grouping_cols: List[str], window: int,
                 standardize: Optional[bool] = False, device=None):
    """Rolling 5-95 % min-max of the centre point (backend/sglm_pp.py:522-545).  device=True (or
    SGLM_DEVICE_DESIGN=1): the rolling windows are evaluated on the GPU (sglm_rolling_minmax_f64: one sorted window
    per position instead of a Python lambda with two pandas quantiles per position) and the same Series comes back
    (group keys + original index as a MultiIndex when `grouping_cols` is given, exactly as pandas builds it)."""
    if device is None:
        device = _device_default()
    if not device:
        target = X.groupby(grouping_cols)[detrend_col] if grouping_cols else X[detrend_col]
        return target.rolling(window=window * 2, center=True).apply(lambda_min_max)
    return _detrend_device(X, detrend_col, grouping_cols, window)


def rolling_minmax_device(x, window, seg_lo=None, seg_hi=None):
    """The rolling statistic of `detrend_data` for a 1-D CUDA tensor x (window = full window length)."""
    torch = nat.require_cuda()
    xd = eng.device_vector(x)
    out = torch.empty_like(xd)
    nat.call("sglm_rolling_minmax_f64", nat.ptr(xd), xd.numel(), nat.ptr(seg_lo), nat.ptr(seg_hi), int(window), nat.ptr(out),
             nat.stream_ptr())
    return out


def _detrend_device(X, detrend_col, grouping_cols, window):
    vals = np.ascontiguousarray(X[detrend_col].to_numpy(dtype=np.float64))
    W = int(window) * 2
    if not grouping_cols:
        out = rolling_minmax_device(vals, W).cpu().numpy()
        return pd.Series(out, index=X.index, name=detrend_col)
    # pandas orders the result by group key, rows of a group in their original order
    grouped = X.groupby(grouping_cols)[detrend_col]
    order = np.concatenate([np.asarray(ix) for ix in grouped.indices.values()]) if len(X) else np.zeros(0, np.int64)
    sizes = [len(ix) for ix in grouped.indices.values()]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    seg_lo = np.repeat(starts[:-1], sizes).astype(np.int64)
    seg_hi = np.repeat(starts[1:], sizes).astype(np.int64)
    out = rolling_minmax_device(vals[order], W, eng._dev(seg_lo, np.int64), eng._dev(seg_hi, np.int64)).cpu().numpy()
    ref_index = grouped.rolling(window=2, center=True).count().index       # the MultiIndex pandas builds (cheap)
    return pd.Series(out, index=ref_index, name=detrend_col)
