"""
sglm_cv — drop-in for the reference module `backend/sglm_cv.py`: cross-validated grid of
penalised GLM fits.  Same entry points and result dictionaries; the execution plan is
B200-native (see _engine.py): the reference runs `len(glm_kwarg_lst) * (n_folds + 1)`
scikit-learn fits one after another from Python threads, copying X[idx_train,:] for
every fold (backend/sglm_cv.py:106-110, :162-170, :180-181); here the whole grid is

    statistics (one pass over X per row set)  ->  ONE batched coordinate-descent launch
    for every ElasticNet/Lasso model + one Cholesky launch per fold for Ridge/OLS  ->
    scores from the statistics,

with identical results (coefficients, per-fold scores, pooled R^2 / MSE, selection).

Reference semantics kept on purpose (results-affecting): `roll` and `model_name` are
popped from the caller's dicts (:95, :288); the full-data refit uses the UN-rolled y
(:181); selection uses strict '>' in list order (:402-415); Ridge alpha is not scaled by
n while ElasticNet's is (backend/sglm.py:102-110).
Not reproduced (crash / hang / waste): the worker-queue hang with fast fits (:22-30,
:169-170) and the discarded PCA + OLS prefit (:275-282).
"""
import itertools

import numpy as np

import _engine as eng
import sglm_


class SGLM_worker():
    """Kept for API compatibility (backend/sglm_cv.py:15-40).  The GPU plan batches the
    work of all workers into single launches, so the queue is drained synchronously."""

    def __init__(self, queue, verbose=0):
        self.queue = queue
        self.verbose = verbose

    def run_single(self):
        while not self.queue.empty():
            glm, args, kwargs = self.queue.get()
            glm.fit_set(*args, **kwargs)
            self.queue.task_done()

    def run_multi(self):
        while not self.queue.empty():
            args, kwargs = self.queue.get()
            cv_glm_single_params(*args, **kwargs)
            self.queue.task_done()


# --------------------------------------------------------------------------- #
# the batched plan
# --------------------------------------------------------------------------- #
def _fold_weights(cv_idx, T):
    """Row multiplicities of every train / test index list (X[idx,:] semantics) and whether
    each train set is exactly the complement of its test set (then train = full - test)."""
    import torch
    te = [eng.index_counts(test, T) for (_, test) in cv_idx]
    tr = [eng.index_counts(train, T) for (train, _) in cv_idx]
    n_te = [int(len(test)) for (_, test) in cv_idx]
    n_tr = [int(len(train)) for (train, _) in cv_idx]
    if cv_idx:
        comp = torch.stack([((a + b) == 1.0).all() for a, b in zip(tr, te)]).cpu().numpy()
    else:
        comp = np.zeros(0, dtype=bool)
    return tr, te, n_tr, n_te, comp


def _use_tensor_core_gram(T, C):
    """SGLM_GRAM=dmma|tc overrides; by default the tensor-core path is used once the Gram is
    large enough for its fixed costs (two analysis passes, plan sync) to pay off."""
    import os
    mode = os.environ.get("SGLM_GRAM", "auto")
    if mode == "dmma":
        return False
    if mode == "tc":
        return True
    return T * C * C >= (1 << 33)


def _unique_sorted_rows(cv_idx, T):
    """Sorted, duplicate-free test rows per fold as CUDA int64 tensors, or None when a test
    list repeats rows (then the weighted fp64 path is used)."""
    import torch
    out = []
    for (_, test) in cv_idx:
        t = test if eng.is_torch(test) else torch.from_numpy(np.ascontiguousarray(np.asarray(test).reshape(-1), dtype=np.int64))
        t = t.to(device="cuda", dtype=torch.int64)
        t = torch.where(t < 0, t + T, t)
        u = torch.unique(t, sorted=True)
        if u.numel() != t.numel():
            return None
        out.append(u)
    return out


def _gaussian_grid(Xd, yd, cv_idx, glms, rolls, score_method):
    """Fit every (param set, fold) + every full-data refit of Gaussian-family GLMs.
    glms: list of sglm_.GLM objects (already dispatched to an estimator class)."""
    import torch
    T, C = Xd.shape
    F = len(cv_idx)
    # y columns: column 0 = un-rolled y (refit, backend/sglm_cv.py:181), then one per distinct roll
    roll_vals = [0] + sorted({r for r in rolls if r % max(T, 1) != 0})
    ycol_of_roll = {}
    cols = []
    for k, r in enumerate(roll_vals):
        cols.append(yd if k == 0 else torch.roll(yd, int(r)))
        ycol_of_roll[r] = k
    for r in rolls:
        if r % max(T, 1) == 0:
            ycol_of_roll[r] = 0
    Yd = torch.stack(cols, dim=1).contiguous()
    n_y = Yd.shape[1]

    tr_w, te_w, n_tr, n_te, comp = _fold_weights(cv_idx, T)
    # statistics: set 0 = full data, 1..F = test folds, then explicit train sets where the
    # train rows are not the complement of the test rows
    extra = [f for f in range(F) if not comp[f]]
    G = None
    if not extra and _use_tensor_core_gram(T, C):
        # 0/1 row sets: tcgen05 int8 digit-plane Gram (exact integer accumulation, fp64 result)
        test_rows = _unique_sorted_rows(cv_idx, T)
        if test_rows is not None:
            G, _ = eng.suffstats_tc(Xd, Yd, [None] + test_rows)
    if G is None:
        # general row multiplicities: fp64 DMMA Gram with row weights
        w_rows = [torch.ones_like(yd)] + te_w + [tr_w[f] for f in extra]
        rows_hint = [T] + n_te + [n_tr[f] for f in extra]
        W = torch.stack(w_rows).contiguous() if len(w_rows) > 1 else None
        G = eng.suffstats(Xd, Yd, W, rows_hint)
        del w_rows, W
    finite_flag = torch.isfinite(G[0]).all()            # read back together with the results
    train_set = {f: (1 + F + extra.index(f)) if f in extra else None for f in range(F)}
    del tr_w

    # centred problems, shared by every model with the same (row set, y column, intercept)
    problems = {}

    def problem(fold, ycol, fi):
        key = (fold, ycol, bool(fi))
        if key not in problems:
            if fold is None:
                p = eng.center(G[0], None, C, n_y, ycol, fi, n_rows=T)
            elif train_set[fold] is None:
                p = eng.center(G[0], G[1 + fold], C, n_y, ycol, fi, n_rows=T - n_te[fold])
            else:
                p = eng.center(G[train_set[fold]], None, C, n_y, ycol, fi, n_rows=n_tr[fold])
            problems[key] = p
        return problems[key]

    specs, owner = [], []            # owner[i] = (param set k, fold f or None)
    for k, (glm, r) in enumerate(zip(glms, rolls)):
        est = glm.model
        est._check_supported()
        for f in list(range(F)) + [None]:
            p = problem(f, ycol_of_roll[r] if f is not None else 0, est.fit_intercept)
            specs.append((est, p))
            owner.append((k, f))
    models = [est._spec(p) for est, p in specs]         # no device read-back: row counts are known on the host
    Wd, info, status = eng.solve_models(models, C)
    b_d, V = eng.finalize(Wd, C, n_y, models)

    # residual sums of squares from the statistics: RSS(set) = V' G[set] V
    M = len(models)
    fold_of = np.array([-1 if f is None else f for (_, f) in owner])
    rss_full = eng.quadform(G[0], V)
    rss_test = torch.zeros(M, dtype=torch.float64, device="cuda")
    rss_train = rss_full.clone()
    for f in range(F):
        sel = np.flatnonzero(fold_of == f)
        if len(sel) == 0:
            continue
        sel_t = eng._dev(sel, np.int64)
        Vf = V.index_select(0, sel_t)
        q_te = eng.quadform(G[1 + f], Vf)
        rss_test.index_copy_(0, sel_t, q_te)
        if train_set[f] is None:
            rss_train.index_copy_(0, sel_t, rss_full.index_select(0, sel_t) - q_te)
        else:
            rss_train.index_copy_(0, sel_t, eng.quadform(G[train_set[f]], Vf))

    if not bool(finite_flag.item()):
        raise ValueError("Input contains NaN, infinity or a value too large for dtype('float64').")
    n_sets_k = len(glms)
    # coefficients leave the device already in the layout of the result dicts: [set][C][F] for the folds
    # (cv_coefs is C x F, backend/sglm_cv.py:98) and [set][C] for the refits — one transposition on the
    # device instead of one per parameter set on the host
    W3 = Wd[:, :C].reshape(n_sets_k, F + 1, C)
    coef_folds = W3[:, :F, :].permute(0, 2, 1).contiguous().cpu().numpy()      # [set][C][F]
    coef_full = W3[:, F, :].contiguous().cpu().numpy()                          # [set][C]
    icpt = b_d.cpu().numpy()
    rss_test_h = rss_test.cpu().numpy()
    rss_train_h = rss_train.cpu().numpy()
    Gy = G[:, C:, C:C + n_y + 1].cpu().numpy()      # [set, (y cols | 1), (y cols | 1)]

    def set_moments(s, ycol):
        n = Gy[s, n_y, n_y]
        sy = Gy[s, ycol, n_y]
        yy = Gy[s, ycol, ycol]
        return n, sy, yy

    # per-fold moments of every y column (vectorised over the folds; the loop below runs once per set)
    mom = {}
    for ycol in sorted(set(ycol_of_roll.values())):
        n_f = np.array([set_moments(1 + f, ycol)[0] for f in range(F)])
        sy_f = np.array([set_moments(1 + f, ycol)[1] for f in range(F)])
        yy_f = np.array([set_moments(1 + f, ycol)[2] for f in range(F)])
        n_t, sy_t, yy_t = np.zeros(F), np.zeros(F), np.zeros(F)
        n0, sy0, yy0 = set_moments(0, ycol)
        for f in range(F):
            if train_set[f] is None:
                n_t[f], sy_t[f], yy_t[f] = n0 - n_f[f], sy0 - sy_f[f], yy0 - yy_f[f]
            else:
                n_t[f], sy_t[f], yy_t[f] = set_moments(train_set[f], ycol)
        with np.errstate(divide='ignore', invalid='ignore'):
            tss_f = np.where(n_f > 0, yy_f - sy_f * sy_f / n_f, 0.0)
            tss_t = np.where(n_t > 0, yy_t - sy_t * sy_t / n_t, 0.0)
        mom[ycol] = (n_f, n_t, tss_f, tss_t)

    def r2_vec(rss, tss):
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.where(tss <= 0.0, np.where(rss == 0.0, 1.0, 0.0), 1.0 - rss / np.where(tss <= 0.0, 1.0, tss))

    # scores of every (set, fold) at once
    ycols = np.array([ycol_of_roll[r] for r in rolls], dtype=np.int64)
    fi = np.array([bool(g.model.fit_intercept) for g in glms])
    rte = np.maximum(rss_test_h.reshape(n_sets_k, F + 1)[:, :F], 0.0)
    rtr = np.maximum(rss_train_h.reshape(n_sets_k, F + 1)[:, :F], 0.0)
    icpt2 = icpt.reshape(n_sets_k, F + 1)
    n_f = np.stack([mom[c][0] for c in ycols]) if n_sets_k else np.zeros((0, F))
    n_t = np.stack([mom[c][1] for c in ycols]) if n_sets_k else np.zeros((0, F))
    tss_f = np.stack([mom[c][2] for c in ycols]) if n_sets_k else np.zeros((0, F))
    tss_t = np.stack([mom[c][3] for c in ycols]) if n_sets_k else np.zeros((0, F))
    if score_method == 'r2':
        S_tr, S_te = r2_vec(rtr, tss_t), r2_vec(rte, tss_f)
    else:
        with np.errstate(divide='ignore', invalid='ignore'):
            S_tr, S_te = -rtr / n_t, -rte / n_f
    rss_pool_a = np.zeros(n_sets_k)
    tss_pool_a = np.zeros(n_sets_k)
    n_pool_a = np.zeros(n_sets_k)
    for f in range(F):                                   # left-to-right, as the reference accumulates fold by fold
        rss_pool_a += rte[:, f]
        tss_pool_a += np.maximum(tss_f[:, f], 0.0)
        n_pool_a += n_f[:, f]
    with np.errstate(divide='ignore', invalid='ignore'):
        mean_tr = S_tr.mean(axis=1) if F else np.full(n_sets_k, np.nan)
        mean_te = S_te.mean(axis=1) if F else np.full(n_sets_k, np.nan)
        std_te = S_te.std(axis=1) if F else np.full(n_sets_k, np.nan)

    results = []
    for k, glm in enumerate(glms):
        base = k * (F + 1)
        rss_pool, tss_pool, n_pool = rss_pool_a[k], tss_pool_a[k], n_pool_a[k]
        i_full = base + F
        _set_fitted(glm, coef_full[k], icpt[i_full], info[i_full], status[i_full])
        results.append({
            'cv_coefs': coef_folds[k],
            'cv_intercepts': icpt2[k, :F].copy() if fi[k] else np.zeros(F),
            'cv_scores_train': S_tr[k].copy(),
            'cv_scores_test': S_te[k].copy(),
            'cv_mean_score_train': mean_tr[k],
            'cv_mean_score': mean_te[k],
            'cv_std_score': std_te[k],
            'cv_R2_score': 0 if tss_pool == 0 else 1 - rss_pool / tss_pool,      # sglm.calc_R2
            'cv_mse_score': rss_pool / n_pool if n_pool else np.nan,
            'model': glm,
            '_fit_info': {'cd_info': info[base:base + F + 1], 'status': status[base:base + F + 1]},
        })
        bad = status[base:base + F + 1]
        if np.any(bad == 2):
            raise np.linalg.LinAlgError("Matrix is singular: X'X + alpha*I is not positive definite")
    return results


def _r2(rss, tss):
    if tss <= 0.0:
        return 1.0 if rss == 0.0 else 0.0
    return 1.0 - rss / tss


def _set_fitted(glm, coef, intercept, info, status):
    est = glm.model
    est.coef_ = np.array(coef, dtype=np.float64)
    est.intercept_ = float(intercept) if est.fit_intercept else 0.0
    est.n_features_in_ = len(coef)
    if est.kind in ("enet", "lasso"):
        est.dual_gap_ = float(info[0])
        est.n_iter_ = int(info[2])
    glm.coef_ = est.coef_
    glm.beta_ = glm.coef_
    glm.intercept_ = est.intercept_
    glm.beta0_ = glm.intercept_


def _poisson_grid(Xd, yd, cv_idx, glms, rolls, score_method):
    """Poisson family: one IRLS solve per (param set, fold) on fold row-weights."""
    import torch
    T, C = Xd.shape
    F = len(cv_idx)
    _, te_w, _, n_te, _ = _fold_weights(cv_idx, T)
    tr_w = [eng.index_counts(train, T) for (train, _) in cv_idx]
    results = []
    # Every IRLS run is driven to the optimum (Newton step below 1e-8), so the starting point only changes the
    # number of iterations: the refit starts from the previous parameter set's refit, the folds from the refit
    # of their own parameter set (a few Newton steps instead of a cold start from log(mean y)).
    prev_full = None
    for glm, r in zip(glms, rolls):
        est = glm.model
        est._check_family()
        y_r = torch.roll(yd, int(r)) if r else yd
        cv_coefs = np.zeros((C, F))
        cv_intercepts = np.zeros(F)
        s_tr, s_te = np.zeros(F), np.zeros(F)
        rss_pool = tss_pool = n_pool = 0.0
        init = prev_full if (prev_full is not None and prev_full[2] == bool(est.fit_intercept)) else (None, None, None)
        w_full, b_full, n_it_full = eng.poisson_irls(Xd, yd, est.alpha, est.fit_intercept, None, est.max_iter, est.tol,
                                                     coef_init=init[0], intercept_init=init[1])
        prev_full = (w_full, b_full, bool(est.fit_intercept))
        for f in range(F):
            w, b, _ = eng.poisson_irls(Xd, y_r, est.alpha, est.fit_intercept, tr_w[f], est.max_iter, est.tol,
                                       coef_init=w_full if not r else None, intercept_init=b_full if not r else None)
            cv_coefs[:, f], cv_intercepts[f] = w, b
            st, _ = eng.score_sums(Xd, y_r, w, b, 1, rw=tr_w[f])
            se, _ = eng.score_sums(Xd, y_r, w, b, 1, rw=te_w[f])
            if score_method == 'r2':
                s_tr[f], s_te[f] = eng.poisson_d2_from_sums(st), eng.poisson_d2_from_sums(se)
            else:
                s_tr[f], s_te[f] = -st[1] / st[0], -se[1] / se[0]
            rss_pool += se[1]
            tss_pool += se[3] - se[2] * se[2] / se[0]
            n_pool += se[0]
        w, b, n_it = w_full, b_full, n_it_full
        est.coef_, est.intercept_, est.n_iter_ = w, b, n_it
        glm.coef_ = glm.beta_ = w
        glm.intercept_ = glm.beta0_ = b
        results.append({
            'cv_coefs': cv_coefs, 'cv_intercepts': cv_intercepts,
            'cv_scores_train': s_tr, 'cv_scores_test': s_te,
            'cv_mean_score_train': np.mean(s_tr), 'cv_mean_score': np.mean(s_te),
            'cv_std_score': np.std(s_te),
            'cv_R2_score': 0 if tss_pool == 0 else 1 - rss_pool / tss_pool,
            'cv_mse_score': rss_pool / n_pool if n_pool else np.nan,
            'model': glm,
        })
    return results


def _cv_batch(X, y, cv_idx, entries, beta_, beta0_, score_method):
    Xd = eng.device_matrix(X)
    yd = eng.device_vector(y)
    if Xd.shape[0] != yd.shape[0]:
        raise ValueError(f"Found input variables with inconsistent numbers of samples: [{Xd.shape[0]}, {yd.shape[0]}]")
    cv_idx = list(cv_idx)
    glms, rolls = [], []
    for model_name, kw in entries:
        rolls.append(int(kw.pop('roll', 0)))                          # backend/sglm_cv.py:95
        glms.append(sglm_.GLM(model_name, beta0_=beta0_, beta_=beta_, **kw))   # refit object (:180)
    out = [None] * len(entries)
    gauss = [i for i, g in enumerate(glms) if g.model.kind in ("ols", "ridge", "lasso", "enet")]
    pois = [i for i, g in enumerate(glms) if g.model.kind == "poisson"]
    if gauss:
        res = _gaussian_grid(Xd, yd, cv_idx, [glms[i] for i in gauss], [rolls[i] for i in gauss], score_method)
        for i, r in zip(gauss, res):
            out[i] = r
    if pois:
        res = _poisson_grid(Xd, yd, cv_idx, [glms[i] for i in pois], [rolls[i] for i in pois], score_method)
        for i, r in zip(pois, res):
            out[i] = r
    for (model_name, kw), r in zip(entries, out):
        r['glm_kwargs'] = kw
    return out


def _order_result(r):
    keys = ['cv_coefs', 'cv_intercepts', 'cv_scores_train', 'cv_scores_test', 'cv_mean_score_train',
            'cv_mean_score', 'cv_std_score', 'cv_R2_score', 'cv_mse_score', 'glm_kwargs', 'model']
    out = {k: r[k] for k in keys}
    for k in r:
        if k not in out:
            out[k] = r[k]
    return out


def cv_glm_single_params(X, y, cv_idx, model_name, glm_kwargs, verbose=0, resp_list=[], beta_=None, beta0_=None,
                         score_method='mse'):
    """Cross-validation for one parameter set: F fold fits + one full-data refit.
    Returns the reference's dict (backend/sglm_cv.py:188-200) and appends it to resp_list."""
    ret_dict = _order_result(_cv_batch(X, y, cv_idx, [(model_name, glm_kwargs)], beta_, beta0_, score_method)[0])
    if verbose > 0:
        print('Completing arguments:', glm_kwargs)
        print(f"{glm_kwargs}\n> cv_mean_score_train: {ret_dict['cv_mean_score_train']}\n> cv_R2_score: "
              f"{ret_dict['cv_R2_score']}\n> cv_mean_score: {ret_dict['cv_mean_score']}")
    resp_list.append(ret_dict)
    return ret_dict


def cv_glm_mult_params(X, y, cv_idx, model_name, glm_kwarg_lst, verbose=0, score_method='mse'):
    """Cross-validation over a list of parameter sets; picks the best by `score_method`
    ('r2' -> pooled cv_R2_score, 'mse' -> cv_mean_score) with the reference's strict '>' in
    list order (backend/sglm_cv.py:210-428)."""
    entries = []
    for glm_kwargs in glm_kwarg_lst:
        if verbose > 0:
            print(glm_kwargs)
        name = glm_kwargs.pop('model_name', 'Gaussian')               # backend/sglm_cv.py:288
        entries.append((name, glm_kwargs))
    resp = [_order_result(r) for r in _cv_batch(X, y, cv_idx, entries, None, None, score_method)]
    if verbose > 0:
        for r in resp:
            print(f"{r['glm_kwargs']}\n> cv_mean_score_train: {r['cv_mean_score_train']}\n> cv_R2_score: "
                  f"{r['cv_R2_score']}\n> cv_mean_score: {r['cv_mean_score']}")
    best_score = -np.inf
    best_params = None
    for cv_result in resp:
        if (score_method == 'r2' and cv_result['cv_R2_score'] > best_score):
            best_score = cv_result['cv_R2_score']
            best_score_std = cv_result['cv_std_score']
            best_params = cv_result['glm_kwargs']
            best_model = cv_result['model']
        elif (score_method == 'mse' and cv_result['cv_mean_score'] > best_score):
            best_score = cv_result['cv_mean_score']
            best_score_std = cv_result['cv_std_score']
            best_params = cv_result['glm_kwargs']
            best_model = cv_result['model']
    return {
        'best_score': best_score,
        'best_score_std': best_score_std,
        'best_params': best_params,
        'best_model': best_model,
        'full_cv_results': resp,
    }


def generate_mult_params(kwarg_lists, kwargs=None):
    """Every combination of the listed values merged over the fixed kwargs: fixed keys first,
    then `kwarg_lists` keys in insertion order, last key varying fastest
    (backend/sglm_cv.py:476-496)."""
    keys = list(kwarg_lists)
    combos = itertools.product(*[list(kwarg_lists[k]) for k in keys])
    out = []
    for combo in combos:
        d = dict(kwargs) if kwargs else {}
        for k, v in zip(keys, combo):
            d[k] = v
        out.append(d)
    return out
