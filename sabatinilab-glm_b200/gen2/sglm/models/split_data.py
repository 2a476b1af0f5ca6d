"""sglm.models.split_data — reference sglm/sglm/models/split_data.py: holdout_split_by_trial_id (:5-100),
holdout_splits (:102-118), cv_idx_by_trial_id (:121-155), cv_idx_from_bucket_ids (:158-188).  Index producers: the
group-level random draws happen on the host with the global numpy RNG exactly as in the reference; with
device="cuda" the expansion to row indices happens on the GPU."""
import numpy as np
import pandas as pd

import sglm_pp


def _bucket_codes(X, id_cols):
    """length-prefixed, '__'-joined id strings -> category codes (split_data.py:36-42; the first generation joins
    with '_' and prefixes only the first column, backend/sglm_ez.py:257-263)."""
    ids = None
    for i, col in enumerate(id_cols):
        txt = X[col].apply(str)
        part = txt.str.len().apply(str) + ':' + txt
        ids = part if i == 0 else ids + '__' + part
    return ids.astype("category").cat.codes


def holdout_split_by_trial_id(X, y=None, id_cols=['nTrial_filenum', 'iBlock'], strat_col=None, strat_mode=None,
                              perc_holdout=0.2):
    """Boolean Series marking held-out rows; groups drawn WITHOUT replacement (split_data.py:99, unlike the first
    generation)."""
    assert len(X) > 0
    bucket_ids = _bucket_codes(X, id_cols)
    n_ids = int(bucket_ids.max() + 1)
    if strat_col is None:
        test_ids = np.random.choice(n_ids, size=int(n_ids * perc_holdout), replace=False)
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


def holdout_splits(dfrel_setup, id_cols=['nTrial_filenum'], perc_holdout=0.2):
    """(setup rows, holdout rows, holdout mask)  (split_data.py:102-118)."""
    holdout = holdout_split_by_trial_id(dfrel_setup, id_cols=id_cols, perc_holdout=perc_holdout)
    return dfrel_setup.loc[~holdout], dfrel_setup.loc[holdout], holdout


def cv_idx_by_trial_id(X, y=None, trial_id_columns=[], num_folds=5, test_size=None, device=None):
    X = pd.DataFrame(X)
    return cv_idx_from_bucket_ids(np.asarray(_bucket_codes(X, trial_id_columns)), X, y=y, num_folds=num_folds,
                                  test_size=test_size, device=device)


def cv_idx_from_bucket_ids(bucket_ids, X, y=None, num_folds=None, test_size=None, device=None):
    return sglm_pp.cv_idx_from_bucket_ids(bucket_ids, X, y=y, num_folds=num_folds, test_size=test_size, device=device)
