"""
sglm_ez — drop-in for the entry points of the reference façade `backend/sglm_ez.py`:
named-column time shifts, CV index builders, `fit_GLM`, `simple_cv_fit`,
`training_fit_holdout_score`.  These are thin adaptors (DataFrame -> values) over
sglm_pp / sglm_cv / sglm_, which run on the B200; the
split / hold-out helpers are CPU logic with the reference's semantics.
Plot and print helpers of the reference module are out of scope.
"""
import numpy as np
import pandas as pd

import sglm_
import sglm_cv
import sglm_pp


def _shift_list(neg_order, pos_order):
    return [0] + list(range(neg_order, 0)) + list(range(1, pos_order + 1))


def timeshift_cols(X, cols_to_shift, neg_order=0, pos_order=1, device=None):
    """All shifts neg_order..pos_order of the named columns, un-shifted frame first
    (backend/sglm_ez.py:102-123).  device=True (extension): a device-resident `sglm_pp.DeviceDesign`
    instead of a host DataFrame — `dropna()`, column selection and the fits then run without the design
    ever crossing PCIe."""
    col_nums = sglm_pp.get_column_nums(X, cols_to_shift)
    return sglm_pp.timeshift_multiple(X, shift_inx=col_nums, shift_amt_list=_shift_list(neg_order, pos_order),
                                      device=device)


def add_timeshifts_to_col_list(all_cols, shifted_cols, neg_order=0, pos_order=1):
    """Column names produced by `timeshift_cols`, shift-major (backend/sglm_ez.py:126-147)."""
    added = []
    for shift_amt in _shift_list(neg_order, pos_order)[1:]:
        added.extend(f'{c}_{shift_amt}' for c in shifted_cols)
    return all_cols + added


def timeshift_cols_by_signal_length(X, cols_to_shift, neg_order=0, pos_order=1, trial_id='nTrial',
                                    dummy_col='nothing', shift_amt_ratio=2.0):
    """Shift each column in steps of (shortest event length // shift_amt_ratio)
    (backend/sglm_ez.py:14-66)."""
    X = X.copy()
    created = dummy_col not in X.columns
    if created:
        X[dummy_col] = 1
    sft_orders = {}
    for col in cols_to_shift:
        shortest = X.query(f'{col} > 0').groupby([trial_id, col])[dummy_col].count().min()
        step = max(shortest // shift_amt_ratio, 1)
        print(f'mnts: {shortest}, sar: {shift_amt_ratio}')
        neg = list(np.arange(neg_order, 0, step))
        pos = list(np.arange(step, pos_order + 1, step))
        sft_orders[col] = (neg, pos)
        X = sglm_pp.timeshift_multiple(X, shift_inx=sglm_pp.get_column_nums(X, [col]),
                                       shift_amt_list=[0] + neg + pos)
    if created:
        X = X.drop(dummy_col, axis=1)
    return X, sft_orders


def add_timeshifts_by_sl_to_col_list(all_cols, shifted_cols, sft_orders):
    added = []
    for col in shifted_cols:
        neg, pos = sft_orders[col]
        # same naming rule as sglm_pp.timeshift_multiple: integral shift amounts are written as integers
        added.extend(col + f'_{int(s) if float(s).is_integer() else s}' for s in neg + pos)
    return all_cols + added


def fit_GLM(X, y, model_name='Gaussian', *args, **kwargs):
    """Fit one GLM on DataFrame X / Series y (backend/sglm_ez.py:149-171)."""
    glm = sglm_.GLM(model_name, *args, **kwargs)
    glm.fit(_vals(X), _vals(y))
    return glm


def _vals(x):
    """`.values` of a DataFrame / Series (backend/sglm_ez.py:376-377); a device-resident design stays as it is."""
    return x if isinstance(x, sglm_pp.DeviceDesign) else x.values


def diff_cols(X, cols, append_to_base=True):
    return sglm_pp.diff(X, sglm_pp.get_column_nums(X, cols), append_to_base=append_to_base)


def cv_idx_by_timeframe(X, y=None, timesteps_per_bucket=20, num_folds=10, test_size=None, device=None):
    """GroupShuffleSplit over time buckets (backend/sglm_ez.py:193-217).  device="cuda": the same splits as
    CUDA index tensors (see sglm_pp.cv_idx_from_bucket_ids)."""
    bucket_ids = sglm_pp.bucket_ids_by_timeframe(X.shape[0], timesteps_per_bucket=timesteps_per_bucket)
    return sglm_pp.cv_idx_from_bucket_ids(bucket_ids, X, y=y, num_folds=num_folds, test_size=test_size, device=device)


def _bucket_codes(X, id_cols):
    ids = None
    for i, col in enumerate(id_cols):
        txt = X[col].astype(str)
        ids = (txt.str.len().astype(str) + ':' + txt) if i == 0 else (ids + '_' + txt)
    return ids.astype("category").cat.codes


def cv_idx_by_trial_id(X, y=None, trial_id_columns=[], num_folds=5, test_size=None, device=None):
    """GroupShuffleSplit keeping trials together (backend/sglm_ez.py:311-343).  device="cuda": the same splits as
    CUDA index tensors."""
    X = pd.DataFrame(X)
    return sglm_pp.cv_idx_from_bucket_ids(np.asarray(_bucket_codes(X, trial_id_columns)), X, y=y, num_folds=num_folds,
                                          test_size=test_size, device=device)


def holdout_split_by_trial_id(X, y=None, id_cols=['nTrial', 'iBlock'], strat_col=None, strat_mode=None,
                              perc_holdout=0.2):
    """Boolean Series marking held-out rows, drawn per trial id with the global numpy RNG
    (backend/sglm_ez.py:219-309)."""
    bucket_ids = _bucket_codes(X, id_cols)
    n_ids = int(bucket_ids.max() + 1)
    if strat_col is None:
        test_ids = np.random.choice(n_ids, size=int(n_ids * perc_holdout))
        return bucket_ids.isin(test_ids)
    frame = X[[strat_col]].copy()
    frame['bucket_id'] = bucket_ids
    groups = [pd.Series(frame[frame[strat_col] == g]['bucket_id'].unique()) for g in frame[strat_col].unique()]
    smallest = min(len(g) for g in groups)
    picked = []
    for g in groups:
        if strat_mode == 'balanced_train':
            train = np.random.choice(g, int(smallest * (1 - perc_holdout)), replace=False)
            picked.append(g[~g.isin(train)])
        elif strat_mode == 'balanced_test':
            picked.append(np.random.choice(g, int(smallest * perc_holdout), replace=False))
        elif strat_mode == 'stratify':
            picked.append(np.random.choice(g, int(len(g) * perc_holdout), replace=False))
        else:
            raise ValueError(f'Invalid strat_mode: {strat_mode}')
    return bucket_ids.isin(np.concatenate(picked))


def simple_cv_fit(X, y, cv_idx, glm_kwarg_lst, model_type='Normal', verbose=0, score_method='mse'):
    """Grid search by cross-validation; returns (best_score, best_score_std, best_params,
    best_model, cv_results) (backend/sglm_ez.py:347-389)."""
    cv_results = sglm_cv.cv_glm_mult_params(_vals(X), _vals(y), cv_idx, model_type, glm_kwarg_lst,
                                            verbose=verbose, score_method=score_method)
    return (cv_results['best_score'], cv_results['best_score_std'], cv_results['best_params'],
            cv_results['best_model'], cv_results)


def training_fit_holdout_score(X_setup, y_setup, X_holdout, y_holdout, best_params):
    """Refit on the training data with the selected parameters and score the hold-out set
    (backend/sglm_ez.py:631-652)."""
    glm = fit_GLM(X_setup, y_setup, **best_params)
    return glm, glm.r2_score(X_holdout, y_holdout), glm.neg_mse_score(X_holdout, y_holdout)


def holdout_scores(models, X_holdout, y_holdout):
    """R^2 and -MSE of MANY fitted Gaussian-family models on one hold-out set in one pass over it — the per-model
    scoring loop of the drivers (er_refactored_from_scratch_cleanup.py:528-550: `fitted_model.r2_score(X_holdout,
    y_holdout)` for every model of `full_cv_results`) as batched score-from-statistics: one Gram of [X | y | 1] of
    the hold-out rows, then RSS_m = v_m' G v_m for all models.  `models`: GLM objects (or anything
    with coef_ / intercept_).  Returns (r2 [M], neg_mse [M]) numpy arrays equal to the per-model calls."""
    import torch
    import _engine as eng
    Xd, yd = eng.device_matrix(_vals(X_holdout)), eng.device_vector(_vals(y_holdout))
    T, C = Xd.shape
    M = len(models)
    G = eng.suffstats(Xd, yd[:, None].contiguous())[0]
    ldv = (C + 2 + 1) // 2 * 2
    V = np.zeros((M, ldv))
    for i, m in enumerate(models):
        V[i, :C] = -np.asarray(m.coef_, dtype=np.float64).reshape(-1)
        V[i, C] = 1.0
        V[i, C + 1] = -float(np.asarray(m.intercept_).reshape(-1)[0])
    rss = np.maximum(eng.quadform(G, eng._dev(V, np.float64)).cpu().numpy(), 0.0)
    mom = G[C:C + 2, C:C + 2].cpu().numpy()                    # [[y'y, sum y], [sum y, n]]
    n, sy, yy = mom[1, 1], mom[0, 1], mom[0, 0]
    tss = yy - sy * sy / n
    with np.errstate(divide='ignore', invalid='ignore'):
        r2 = np.where(tss <= 0.0, np.where(rss == 0.0, 1.0, 0.0), 1.0 - rss / (tss if tss > 0 else 1.0))
    return r2, -rss / n


def calc_l1(coeffs):
    return np.sum(np.abs(coeffs))


def calc_l2(coeffs):
    return np.sum(np.square(coeffs))
