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


def timeshift_multiple(X, shift_inx=[], shift_amt_list=[-1, 0, 1], unshifted_keep_all=True, fill_value=np.nan):
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
        return eng.gather(eng.device_matrix(X), src, sh, fill_value)
    if is_df:
        vals = _gather_host(X.values, src, sh, fill_value)
        names, pos = [], 0
        for a, n in zip(shift_amt_list, sizes):
            base = X.columns[src[pos:pos + n]]
            names.extend(list(base) if a == 0 else [f"{c}_{a}" for c in base])
            pos += n
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
    return (X - X.mean(axis=0)) / X.std(axis=0)


def diff(X, diff_inx=[], n=1, axis=0, append_to_base=False, fill_value=np.nan, **kwargs):
    """n-th difference of the chosen columns (backend/sglm_pp.py:120-190)."""
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
                 standardize: Optional[bool] = False):
    """Rolling 5-95 % min-max of the centre point (backend/sglm_pp.py:522-545)."""
    target = X.groupby(grouping_cols)[detrend_col] if grouping_cols else X[detrend_col]
    return target.rolling(window=window * 2, center=True).apply(lambda_min_max)
