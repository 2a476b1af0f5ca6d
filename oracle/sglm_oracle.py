"""
ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the sGLM hot path (kimerein/sabatinilab-glm, `backend/`):
lag/shift design-matrix construction, the `GLM` estimator dispatch, one fit, the
scores, and the cross-validation grid.  Nothing under `sabatinilab-glm_b200/`
imports this module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs do, as the checker or as the reported
CPU baseline.

Where the arithmetic lives.  The reference delegates every fit to scikit-learn
(`backend/sglm.py:2`, construction at `:101,105,108,110,115`, `fit` at `:241`);
scikit-learn is a third-party dependency that is not vendored in the reference
tree (`requirements.txt:7` pins scikit_learn==0.24.2; this image has 1.9.0).
Two engines are therefore offered:

* ``engine="restated"`` — numpy / plain-C restatement of the published
  algorithms (cyclic coordinate descent `_cd_fast.pyx:243-506`, Ridge Cholesky
  `_ridge.py:215-227`, centred least squares `_base.py:700-756`, Poisson
  deviance + L2 `_glm/glm.py:185-339`).  This is the checker.
* ``engine="sklearn"`` — the same wrapper logic, but handing the fit to the
  installed scikit-learn estimators exactly as `backend/sglm.py:95-130` does.
  This is the faithful CPU baseline (what the reference executes on a host).

Parity pin: tests/test_oracle_pins.py checks both engines against fixtures in
tests/golden/ that scripts/make_golden.py produced by importing the unmodified
reference modules from /root/reference/backend in the build container
(scikit-learn 1.9.0, numpy 2.3.5), and against the seven `test_sglm_pp.py`
cases of the reference's own test-suite (backend/test/test_sglm_pp.py:20-151).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

try:  # pandas is only needed for the DataFrame flavoured helpers
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_LOCK = threading.Lock()


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/enet_cd_oracle.c -> oracle/_build/liboracle.so (gcc)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "enet_cd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, src, "-lm"]
        )
    return so


def _lib():
    global _LIB
    with _LIB_LOCK:
        if _LIB is None:
            lib = ctypes.CDLL(build_c_oracle())
            dp = ctypes.POINTER(ctypes.c_double)
            ip = ctypes.POINTER(ctypes.c_int)
            lib.sglm_oracle_enet_cd.restype = ctypes.c_int
            lib.sglm_oracle_enet_cd.argtypes = [
                dp, ctypes.c_double, ctypes.c_double, dp, dp, ctypes.c_int, ctypes.c_int,
                ctypes.c_int, ctypes.c_double, ctypes.c_int, dp, dp, ip]
            lib.sglm_oracle_enet_cd_gram.restype = ctypes.c_int
            lib.sglm_oracle_enet_cd_gram.argtypes = [
                dp, ctypes.c_double, ctypes.c_double, dp, dp, ctypes.c_double, ctypes.c_int,
                ctypes.c_int, ctypes.c_double, ctypes.c_int, dp, dp, ip,
                ctypes.POINTER(ctypes.c_longlong)]
            lib.sglm_oracle_timeshift.restype = None
            lib.sglm_oracle_timeshift.argtypes = [
                dp, ctypes.c_longlong, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_double, dp]
            _LIB = lib
    return _LIB


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------- #
# (a1-a3) lag / shift gather            backend/sglm_pp.py:23-103, :298-357, :436-486
# --------------------------------------------------------------------------- #
def shift(a, shift_amt, fill_value=np.nan):
    """out[t] = a[t - shift_amt] where that row exists, else fill (sglm_pp.py:298-357).

    A non-zero shift always yields float64 (the blanks block is float64,
    sglm_pp.py:312); a zero shift returns the input object itself (:317-318)."""
    if shift_amt == 0:
        return a
    T = a.shape[0]
    out = np.full((T, a.shape[1]), fill_value, dtype=np.result_type(a.dtype, np.float64))
    k = abs(int(shift_amt))
    if k < T:
        if shift_amt > 0:
            out[k:, :] = a[: T - k, :]
        else:
            out[: T - k, :] = a[k:, :]
    return out


def _is_df(X):
    return pd is not None and type(X) == pd.DataFrame


def timeshift(X, shift_inx=[], shift_amt=1, keep_non_inx=False, dct=None, fill_value=np.nan):
    """One shift block (sglm_pp.py:23-56, :359-434)."""
    vals = X.values if _is_df(X) else X
    cols = list(range(vals.shape[1])) if len(shift_inx) == 0 else list(shift_inx)
    moved = shift(vals[:, cols], shift_amt, fill_value=fill_value)
    if _is_df(X):
        res = X.copy()
        # pandas >= 2 refuses silent int -> float upcasts on iloc assignment; the
        # reference (pandas 1.1.3) upcast.  Build the frame column by column.
        for i, c in enumerate(cols):
            name = res.columns[c]
            res[name] = moved[:, i]
        if not keep_non_inx:
            res = res.iloc[:, cols]
    else:
        if keep_non_inx:
            res = X.copy()          # keeps X's dtype: values are cast on assignment (:430-431)
            res[:, cols] = moved
        else:
            res = moved.copy()
    if dct is not None:
        dct[shift_amt] = res
    return res


def timeshift_multiple(X, shift_inx=[], shift_amt_list=[-1, 0, 1], unshifted_keep_all=True,
                       fill_value=np.nan):
    """Column-concatenated shift blocks in list order (sglm_pp.py:58-103, :436-486)."""
    blocks = {}
    for a in shift_amt_list:
        blocks[a] = timeshift(X, shift_inx=shift_inx, shift_amt=a,
                              keep_non_inx=(a == 0 and unshifted_keep_all), fill_value=fill_value)
    ordered = [blocks[a] for a in shift_amt_list]
    if _is_df(X):
        renamed = []
        for a, blk in zip(shift_amt_list, ordered):
            if a != 0:
                blk = blk.rename({c: f"{c}_{a}" for c in blk.columns}, axis=1)
            renamed.append(blk)
        return pd.concat(renamed, axis=1)
    return np.concatenate(ordered, axis=1)


def column_map(n_cols, shift_inx, shift_amt_list, unshifted_keep_all=True):
    """(col_src, col_shift) of the output of timeshift_multiple — the layout contract
    (shift-major, predictor-minor; the 0-shift block holds all columns)."""
    cols = list(range(n_cols)) if len(shift_inx) == 0 else list(shift_inx)
    src, sh = [], []
    for a in shift_amt_list:
        if a == 0 and unshifted_keep_all:
            src.extend(range(n_cols)); sh.extend([0] * n_cols)
        else:
            src.extend(cols); sh.extend([a] * len(cols))
    return np.asarray(src, dtype=np.int32), np.asarray(sh, dtype=np.int32)


def timeshift_c(X, col_src, col_shift, fill_value=np.nan):
    """C loop version of the gather (timed CPU baseline of the gather only)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    out = np.empty((X.shape[0], len(col_src)), dtype=np.float64)
    cs = np.ascontiguousarray(col_src, dtype=np.int32)
    sh = np.ascontiguousarray(col_shift, dtype=np.int32)
    _lib().sglm_oracle_timeshift(_dptr(X), X.shape[0], X.shape[1],
                                 cs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                 sh.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                 len(cs), float(fill_value), _dptr(out))
    return out


# --------------------------------------------------------------------------- #
# (a6-a9) single fits, restated
# --------------------------------------------------------------------------- #
def _centre(X, y, fit_intercept):
    """sklearn/linear_model/_base.py:_preprocess_data — centre X and y (copy)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    if fit_intercept:
        x_off = X.mean(axis=0)
        y_off = y.mean()
        return X - x_off, y - y_off, x_off, y_off
    return X, y, np.zeros(X.shape[1]), 0.0


def enet_fit(X, y, alpha=1.0, l1_ratio=0.5, fit_intercept=True, max_iter=1000, tol=1e-4,
             coef_init=None, do_screening=True, use_gram=False):
    """ElasticNet / Lasso objective 1/(2n)||y-Xw-b||^2 + a*l1*|w|_1 + a*(1-l1)/2*|w|^2
    (sklearn/linear_model/_coordinate_descent.py:781-782 scaling, :1281 intercept)."""
    Xc, yc, x_off, y_off = _centre(X, y, fit_intercept)
    n, p = Xc.shape
    l1_reg = alpha * l1_ratio * n
    l2_reg = alpha * (1.0 - l1_ratio) * n
    w = np.zeros(p) if coef_init is None else np.array(coef_init, dtype=np.float64)
    gap = ctypes.c_double()
    tol_s = ctypes.c_double()
    n_iter = ctypes.c_int()
    if use_gram:
        Q = np.ascontiguousarray(Xc.T @ Xc)
        q = np.ascontiguousarray(Xc.T @ yc)
        nup = ctypes.c_longlong()
        rc = _lib().sglm_oracle_enet_cd_gram(_dptr(w), l1_reg, l2_reg, _dptr(Q), _dptr(q),
                                             float(yc @ yc), p, int(max_iter), float(tol),
                                             int(do_screening), ctypes.byref(gap),
                                             ctypes.byref(tol_s), ctypes.byref(n_iter),
                                             ctypes.byref(nup))
    else:
        Xf = np.asfortranarray(Xc)
        yc = np.ascontiguousarray(yc)
        rc = _lib().sglm_oracle_enet_cd(_dptr(w), l1_reg, l2_reg, _dptr(Xf), _dptr(yc), n, p,
                                        int(max_iter), float(tol), int(do_screening),
                                        ctypes.byref(gap), ctypes.byref(tol_s),
                                        ctypes.byref(n_iter))
    b = y_off - x_off @ w if fit_intercept else 0.0
    return w, float(b), dict(gap=gap.value, tol=tol_s.value, n_iter=n_iter.value, converged=rc == 0)


def ridge_fit(X, y, alpha=1.0, fit_intercept=True):
    """(Xc'Xc + alpha I) w = Xc'yc — alpha NOT scaled by n (sklearn _ridge.py:215-227)."""
    Xc, yc, x_off, y_off = _centre(X, y, fit_intercept)
    A = Xc.T @ Xc
    A.flat[:: A.shape[0] + 1] += alpha
    w = np.linalg.solve(A, Xc.T @ yc)
    b = y_off - x_off @ w if fit_intercept else 0.0
    return w, float(b)


def ols_fit(X, y, fit_intercept=True, cond=1e-6):
    """Centred minimum-norm least squares (sklearn _base.py:700-756; scipy lstsq with
    cond = LinearRegression.tol = 1e-6 in 1.9.0)."""
    Xc, yc, x_off, y_off = _centre(X, y, fit_intercept)
    w = np.linalg.lstsq(Xc, yc, rcond=cond)[0]
    b = y_off - x_off @ w if fit_intercept else 0.0
    return w, float(b)


def poisson_fit(X, y, alpha=1.0, fit_intercept=True, max_iter=100, tol=1e-10):
    """argmin mean(mu - y*eta) + alpha/2 ||w||^2, eta = Xw + b, mu = exp(eta)
    (sklearn _glm/glm.py:185-339: log link, intercept un-penalised, start at
    w = 0, b = log(mean y)).  Solved to the optimum by Newton with step halving —
    the optimum is what the parity tests compare (sklearn's L-BFGS stops at
    gtol = 1e-4, so fixtures are generated at a tightened tol)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, p = X.shape
    w = np.zeros(p)
    b = float(np.log(y.mean())) if fit_intercept else 0.0

    def objective(w_, b_):
        eta = X @ w_ + b_
        return float(np.mean(np.exp(eta) - y * eta) + 0.5 * alpha * (w_ @ w_))

    f = objective(w, b)
    it = 0
    for it in range(1, max_iter + 1):
        mu = np.exp(X @ w + b)
        g_eta = (mu - y) / n
        gw = X.T @ g_eta + alpha * w
        H = (X.T * (mu / n)) @ X
        H.flat[:: p + 1] += alpha
        if fit_intercept:
            gb = g_eta.sum()
            hb = X.T @ (mu / n)
            Hf = np.empty((p + 1, p + 1))
            Hf[:p, :p] = H
            Hf[:p, p] = hb
            Hf[p, :p] = hb
            Hf[p, p] = mu.sum() / n
            g = np.concatenate([gw, [gb]])
        else:
            Hf, g = H, gw
        if np.max(np.abs(g)) <= tol:
            break
        step = np.linalg.solve(Hf, -g)
        t = 1.0
        while True:
            w_new = w + t * step[:p]
            b_new = b + t * step[p] if fit_intercept else 0.0
            f_new = objective(w_new, b_new)
            if f_new <= f + 1e-4 * t * (g @ step) or t < 1e-10:
                break
            t *= 0.5
        w, b, f = w_new, b_new, f_new
    return w, float(b), dict(n_iter=it)


def r2_score(y, pred):
    """sklearn.metrics.r2_score as used by RegressorMixin.score (sglm.py:184)."""
    y = np.asarray(y, dtype=np.float64)
    rss = np.sum((y - pred) ** 2)
    tss = np.sum((y - y.mean()) ** 2)
    if tss == 0:
        return 1.0 if rss == 0 else 0.0
    return float(1.0 - rss / tss)


def poisson_d2(y, mu):
    """D^2 = 1 - dev(y, mu)/dev(y, mean y) (sklearn _glm/glm.py:387-463)."""
    y = np.asarray(y, dtype=np.float64)

    def dev(m):
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(y > 0, y * np.log(y / m), 0.0)
        return float(np.mean(2.0 * (t - y + m)))

    return 1.0 - dev(mu) / dev(np.full_like(y, y.mean()))


# --------------------------------------------------------------------------- #
# (a5, a10-a12) GLM wrapper                         backend/sglm.py:24-408
# --------------------------------------------------------------------------- #
class GLM:
    """Restatement of `sglm.GLM` (backend/sglm.py:59-347).

    engine="restated": own numerics above.  engine="sklearn": construct the same
    scikit-learn estimator the reference constructs (sglm.py:95-130)."""

    def __init__(self, model_name, beta0_=None, beta_=None, score_method="mse",
                 engine="restated", **kwargs):
        if "warm_start" not in kwargs and (beta0_ is not None or isinstance(beta_, np.ndarray)):
            kwargs["warm_start"] = True
        self.model_name = model_name
        self.engine = engine
        kind = None
        if model_name in {"Normal", "Gaussian"}:
            if "alpha" in kwargs and kwargs["alpha"] == 0:          # sglm.py:96-101
                kwargs.pop("alpha"); kwargs.pop("l1_ratio"); kwargs.pop("max_iter")
                kwargs.pop("warm_start", None)
                kind = "ols"
            elif "l1_ratio" in kwargs and kwargs["l1_ratio"] == 0:  # sglm.py:102-105
                del kwargs["l1_ratio"]; kwargs.pop("warm_start", None)
                kind = "ridge"
            elif "l1_ratio" in kwargs and kwargs["l1_ratio"] == 1:  # sglm.py:106-108
                del kwargs["l1_ratio"]
                kind = "lasso"
            else:                                                    # sglm.py:109-110
                kind = "enet"
        elif model_name == "Poisson":                                # sglm.py:112-115
            kind = "poisson"
        elif model_name in {"PCA Normal", "PCA Gaussian"}:
            kind = "ols"
        else:
            raise NotImplementedError(f"oracle does not restate model_name={model_name!r}")
        self.kind = kind
        self.kwargs = kwargs
        self.beta0_init = beta0_
        self.beta_init = np.copy(beta_) if isinstance(beta_, np.ndarray) else None
        self.score = self.r2_score if score_method == "r2" else self.neg_mse_score
        self.model = None
        if engine == "sklearn":
            from sklearn.linear_model import (ElasticNet, Lasso, LinearRegression, Ridge,
                                              TweedieRegressor)
            if kind == "poisson":
                self.model = TweedieRegressor(power=1, **kwargs)
            else:
                Base = {"ols": LinearRegression, "ridge": Ridge, "lasso": Lasso,
                        "enet": ElasticNet}[kind]
                self.model = Base(**kwargs)
                if beta0_ is not None:
                    self.model.intercept_ = beta0_
                if self.beta_init is not None:
                    self.model.coef_ = self.beta_init

    def fit(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        kw = self.kwargs
        if self.engine == "sklearn":
            self.model.fit(X, y)
            self.coef_ = self.model.coef_
            self.intercept_ = self.model.intercept_
        else:
            fi = kw.get("fit_intercept", True)
            if self.kind == "ols":
                self.coef_, self.intercept_ = ols_fit(X, y, fi)
            elif self.kind == "ridge":
                self.coef_, self.intercept_ = ridge_fit(X, y, kw.get("alpha", 1.0), fi)
            elif self.kind in ("lasso", "enet"):
                l1 = 1.0 if self.kind == "lasso" else kw.get("l1_ratio", 0.5)
                init = self.beta_init if kw.get("warm_start", False) else None
                self.coef_, self.intercept_, self.info_ = enet_fit(
                    X, y, kw.get("alpha", 1.0), l1, fi, kw.get("max_iter", 1000),
                    kw.get("tol", 1e-4), coef_init=init)
            elif self.kind == "poisson":
                self.coef_, self.intercept_, self.info_ = poisson_fit(
                    X, y, kw.get("alpha", 1.0), fi, max_iter=max(100, kw.get("max_iter", 100)),
                    tol=min(1e-10, kw.get("tol", 1e-10)))
        self.beta_ = self.coef_
        self.beta0_ = self.intercept_
        return self

    def predict(self, X):
        if _is_df(X):
            X = X.values
        eta = np.asarray(X, dtype=np.float64) @ self.coef_ + self.intercept_
        return np.exp(eta) if self.kind == "poisson" else eta

    def neg_mse_score(self, X, y):                       # sglm.py:150-167
        r = np.asarray(y, dtype=np.float64) - self.predict(X)
        return -np.mean(r ** 2)

    def r2_score(self, X, y):                            # sglm.py:169-184
        pred = self.predict(X)
        return poisson_d2(y, pred) if self.kind == "poisson" else r2_score(y, pred)

    def get_residuals(self, X, y):                       # sglm.py:314-331
        y = np.asarray(y, dtype=np.float64)
        return y - self.predict(X), y - np.mean(y)


def calc_R2(residuals, mean_residuals):                  # sglm.py:388-408
    rss = np.sum(residuals ** 2)
    tss = np.sum(mean_residuals ** 2)
    return 0 if tss == 0 else 1 - rss / tss


# --------------------------------------------------------------------------- #
# (a13-a15) CV grid                                 backend/sglm_cv.py:42-428, :476-496
# --------------------------------------------------------------------------- #
def cv_glm_single_params(X, y, cv_idx, model_name, glm_kwargs, score_method="mse",
                         engine="restated", n_threads=1):
    """F fold fits + one full-data refit for one parameter set (sglm_cv.py:42-206).
    `roll` is popped from the caller's dict (:95); the refit uses the un-rolled y (:181)."""
    n_coefs, n_idx = X.shape[1], len(cv_idx)
    roll = glm_kwargs.pop("roll", 0)
    y_rolled = np.roll(np.asarray(y).reshape(-1), roll)
    cv_coefs = np.zeros((n_coefs, n_idx))
    cv_intercepts = np.zeros(n_idx)
    tr = np.zeros(n_idx)
    te = np.zeros(n_idx)
    resids = [None] * n_idx
    mean_resids = [None] * n_idx

    def one(k):
        itr, ite = cv_idx[k]
        Xtr, ytr, Xte, yte = X[itr, :], y_rolled[itr], X[ite, :], y_rolled[ite]
        g = GLM(model_name, score_method=score_method, engine=engine, **glm_kwargs)
        g.fit(Xtr, ytr)
        cv_coefs[:, k] = g.coef_
        cv_intercepts[k] = g.intercept_
        tr[k] = g.score(Xtr, ytr)
        te[k] = g.score(Xte, yte)
        resids[k], mean_resids[k] = g.get_residuals(Xte, yte)

    if n_threads > 1:                                    # reference: 4 fold threads (:162-170)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(one, range(n_idx)))
    else:
        for k in range(n_idx):
            one(k)
    full = GLM(model_name, engine=engine, **glm_kwargs)
    full.fit(X, y)
    return {
        "cv_coefs": cv_coefs, "cv_intercepts": cv_intercepts,
        "cv_scores_train": tr, "cv_scores_test": te,
        "cv_mean_score_train": np.mean(tr), "cv_mean_score": np.mean(te),
        "cv_std_score": np.std(te),
        "cv_R2_score": calc_R2(np.concatenate(resids), np.concatenate(mean_resids)),
        "cv_mse_score": np.mean(np.square(np.concatenate(resids))),
        "glm_kwargs": glm_kwargs, "model": full,
    }


def cv_glm_mult_params(X, y, cv_idx, model_name, glm_kwarg_lst, score_method="mse",
                       engine="restated", n_threads=1):
    """Serial loop over parameter sets + strict-'>' selection (sglm_cv.py:210-428).
    The discarded PCA prefit (:275-277) is not restated — its result is thrown away
    (:281-282) and it does not influence any returned value."""
    resp = []
    for kw in glm_kwarg_lst:
        mn = kw.pop("model_name", "Gaussian")            # sglm_cv.py:288
        resp.append(cv_glm_single_params(X, y, cv_idx, mn, kw, score_method, engine, n_threads))
    best_score, best = -np.inf, None
    for r in resp:
        s = r["cv_R2_score"] if score_method == "r2" else r["cv_mean_score"]
        if score_method in ("r2", "mse") and s > best_score:
            best_score, best = s, r
    return {
        "best_score": best_score,
        "best_score_std": best["cv_std_score"],
        "best_params": best["glm_kwargs"],
        "best_model": best["model"],
        "full_cv_results": resp,
    }


def generate_mult_params(kwarg_lists, kwargs=None):
    """Cartesian product; fixed kwargs first, last key fastest (sglm_cv.py:476-496)."""
    import itertools
    keys = list(kwarg_lists)
    out = []
    for combo in itertools.product(*[list(kwarg_lists[k]) for k in keys]):
        d = dict(kwargs) if kwargs else {}
        d.update(dict(zip(keys, combo)))
        out.append(d)
    return out


# synthetic "photometry-shaped" inputs live in the neutral module synth_data.py (repo root)
import sys as _sys
_sys.path.insert(0, os.path.dirname(_HERE))
from synth_data import synth_base, synth_folds, synth_kernels, synth_response  # noqa: E402,F401
