"""
setup_model_fit — the second-generation lag builder of the reference
(`sglm/sglm/features/setup_model_fit.py:43-96`, SURVEY.md §8f "next" row 2) on the same
B200 gather kernel: a second column layout for the same design-matrix construction.

Layout (predictor-major, unlike sglm_pp.timeshift_multiple which is shift-major): the
original frame, then for every column of `X_cols_dict` (insertion order) one column per shift
`neg_order .. pos_order` INCLUDING 0, named f"{col}_{shift}"; rows with NaN in the extreme
shifts are dropped unless `keep_nans`.  Reference quirk kept: the drop list is built from the
(neg_order, pos_order) of the LAST dictionary entry for all columns (:88-93).
"""
import numpy as np
import pandas as pd

import _engine as eng


def timeshift_vals_by_dict(df, X_cols_dict, keep_nans=False):
    src_cols, shifts, names = [], [], []
    neg_order = pos_order = None
    for X_col in X_cols_dict:
        neg_order, pos_order = X_cols_dict[X_col]
        loc = df.columns.get_loc(X_col)
        if not isinstance(loc, (int, np.integer)):
            raise ValueError('Duplicate column found in X column names.')
        for shift_amt in range(neg_order, pos_order + 1):
            src_cols.append(int(loc))
            shifts.append(int(shift_amt))
            names.append(X_col + '_' + str(shift_amt))
    if names:
        cols = sorted(set(src_cols))
        remap = {c: i for i, c in enumerate(cols)}
        base = eng.device_matrix(np.ascontiguousarray(df.iloc[:, cols].to_numpy(dtype=np.float64)))
        vals = eng.gather(base, np.array([remap[c] for c in src_cols], dtype=np.int32),
                          np.array(shifts, dtype=np.int32), np.nan).cpu().numpy()
        shifted = pd.DataFrame(vals, index=df.index, columns=names)
        out = pd.concat([df.copy(), shifted], axis=1)
    else:
        out = df.copy()
    if not keep_nans and X_cols_dict:
        na_drop_cols = [c + '_' + str(neg_order) for c in X_cols_dict] + [c + '_' + str(pos_order) for c in X_cols_dict]
        na_drop_cols = [c for c in na_drop_cols if c in out.columns]
        out = out.dropna(subset=na_drop_cols)
    return out, names


def X_cols_dict_to_default(X_cols_dict, neg_order=-20, pos_order=20):
    """(0, 0) or None entries take the default orders (setup_model_fit.py:98-105)."""
    X_cols_dict = X_cols_dict.copy()
    for X_col in X_cols_dict:
        if X_cols_dict[X_col] == (0, 0) or X_cols_dict[X_col] is None:
            X_cols_dict[X_col] = (neg_order, pos_order)
    return X_cols_dict
