"""
Generate the golden fixtures under tests/golden/ by importing the UNMODIFIED
reference modules from /root/reference/backend (sglm, sglm_pp, sglm_cv) together
with the scikit-learn installed in the build container.  /root/reference does
not exist on the GPU box, so the outputs are committed; this script is the
record of how they were made.  Run from the repo root:

    PYTHONDONTWRITEBYTECODE=1 python scripts/make_golden.py

Versions are stored in every fixture (the reference pins scikit_learn==0.24.2 in
requirements.txt:7; the container has 1.9.0 — same objectives).
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/backend"
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
os.chdir("/tmp")

import pandas as pd  # noqa: E402
import scipy  # noqa: E402
import sklearn  # noqa: E402
import sglm  # noqa: E402  (reference)
import sglm_cv  # noqa: E402  (reference)
import sglm_pp  # noqa: E402  (reference)
from sklearn.linear_model import TweedieRegressor  # noqa: E402

from oracle import sglm_oracle as orc  # noqa: E402  (inputs only: synthetic generators)

# Harness-only guard (no reference source is modified): the reference's fold workers
# return when the queue is empty after a task (backend/sglm_cv.py:22-30); a worker that
# never received a task blocks in Queue.get() forever and thread.join() (:169-170) hangs
# whenever fits are fast.  Give the blocked get() a timeout so that worker ends instead.
import queue as _queue  # noqa: E402
import threading as _threading  # noqa: E402


class _TimedQueue(_queue.Queue):
    def get(self, block=True, timeout=None):
        return super().get(block, 1.0 if timeout is None else timeout)


_queue.Queue = _TimedQueue
_threading.excepthook = lambda args: None

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
VERSIONS = dict(sklearn=sklearn.__version__, numpy=np.__version__, scipy=scipy.__version__,
                pandas=pd.__version__, reference="kimerein/sabatinilab-glm backend/")
warnings.filterwarnings("ignore")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# --------------------------------------------------------------------------- #
# 1. gather: the reference's own unit-test cases + NaN / edge variants
# --------------------------------------------------------------------------- #
def gather_cases():
    base_int = np.arange(20).reshape(5, 4)
    rng = np.random.default_rng(0)
    base_f = rng.standard_normal((37, 6))
    base_f[3, 2] = np.nan            # NaNs already present in the data must survive bit-exactly
    base_f[10, 0] = -0.0
    base_f[11, 1] = np.inf
    cases = []

    def add(name, X, **kw):
        res = sglm_pp.timeshift_multiple(X, **kw) if "shift_amt_list" in kw else sglm_pp.timeshift(X, **kw)
        cases.append((name, X, kw, np.asarray(res)))

    # backend/test/test_sglm_pp.py:20-151 (fill_value=0, 5x4 arange)
    add("unshifted_all", base_int, shift_amt=0)
    add("unshifted_sub", base_int, shift_inx=[0, 3], shift_amt=0)
    add("fwd_all", base_int, shift_amt=1, fill_value=0)
    add("fwd_sub", base_int, shift_inx=[0, 3], shift_amt=1, fill_value=0)
    add("bwd_all", base_int, shift_amt=-1, fill_value=0)
    add("bwd_sub", base_int, shift_inx=[0, 3], shift_amt=-1, fill_value=0)
    add("keep_fwd", base_int, shift_inx=[0, 1], shift_amt=1, fill_value=0, keep_non_inx=True)
    add("keep_bwd", base_int, shift_inx=[0, 1], shift_amt=-1, fill_value=0, keep_non_inx=True)
    add("multi_all", base_int, shift_amt_list=[-1, 0, 1], unshifted_keep_all=True, fill_value=0)
    add("multi_sub", base_int, shift_inx=[0, 3], shift_amt_list=[-1, 0, 1],
        unshifted_keep_all=True, fill_value=0)
    # NaN padding (default fill) — never asserted by the reference tests (NaN != NaN)
    add("nan_fwd", base_f, shift_amt=3)
    add("nan_bwd_sub", base_f, shift_inx=[5, 1], shift_amt=-4)
    add("nan_multi", base_f, shift_inx=[1, 4], shift_amt_list=[0, -3, -2, -1, 1, 2, 3])
    add("nan_multi_nokeep", base_f, shift_inx=[2], shift_amt_list=[2, 0, -2], unshifted_keep_all=False)
    add("nan_multi_zero_not_first", base_f, shift_inx=[0, 5], shift_amt_list=[-1, 1, 0])
    add("big_shift", base_f, shift_amt=40)                # |shift| >= T  -> all fill
    add("big_shift_neg", base_f, shift_inx=[3], shift_amt=-37)
    add("edge_T_minus_1", base_f, shift_amt=36)
    add("one_row", base_f[:1].copy(), shift_amt_list=[-1, 0, 1])
    add("fill_custom", base_f, shift_inx=[0], shift_amt_list=[0, 5, -5], fill_value=-7.25)
    add("repeat_shift", base_f, shift_inx=[1], shift_amt_list=[1, 1, 0])
    return cases


def save_gather():
    cases = gather_cases()
    blob = {}
    meta = []
    for i, (name, X, kw, res) in enumerate(cases):
        blob[f"x{i}"] = X
        blob[f"r{i}"] = res
        meta.append(dict(name=name, kwargs={k: (v if not isinstance(v, float) or v == v else "nan")
                                            for k, v in kw.items()},
                         res_dtype=str(res.dtype)))
    # DataFrame naming contract (test_sglm_pp.py:138-148)
    df = pd.DataFrame(np.arange(20).reshape(5, 4), columns=list("ABCD"))
    names_all = list(sglm_pp.timeshift_multiple(df, shift_amt_list=[-1, 0, 1], fill_value=0).columns)
    names_sub = list(sglm_pp.timeshift_multiple(df, shift_inx=[0, 3], shift_amt_list=[-1, 0, 1],
                                                fill_value=0).columns)
    np.savez_compressed(os.path.join(OUT, "gather_ref.npz"), **blob)
    with open(os.path.join(OUT, "gather_ref.json"), "w") as f:
        json.dump(dict(versions=VERSIONS, cases=meta, df_names_all=names_all,
                       df_names_sub=names_sub), f, indent=1)
    print("gather cases:", len(cases))


# --------------------------------------------------------------------------- #
# 2. single fits through the reference GLM wrapper
# --------------------------------------------------------------------------- #
def small_problem(seed=3, T=3000, P=5, h=4, poisson=False):
    X0 = orc.synth_base(T, P, seed)
    shifts = [0] + list(range(-h, 0)) + list(range(1, h + 1))
    Xd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)
    keep = ~np.isnan(Xd).any(axis=1)
    Xd = Xd[keep]
    beta = orc.synth_kernels(P, shifts, seed)
    y = orc.synth_response(Xd, beta, seed, poisson=poisson)
    return X0, shifts, keep, Xd, y


def save_fits():
    X0, shifts, keep, Xd, y = small_problem()
    grid = []
    for fi in (True, False):
        grid.append(dict(alpha=0, l1_ratio=0, max_iter=1000, fit_intercept=fi))          # OLS
        for a in (1e-3, 1.0, 100.0):
            grid.append(dict(alpha=a, l1_ratio=0, max_iter=1000, fit_intercept=fi))      # Ridge
        for a in (1e-4, 1e-3, 1e-2, 1e-1):
            grid.append(dict(alpha=a, l1_ratio=1, max_iter=1000, fit_intercept=fi))      # Lasso
            for l1 in (0.1, 0.5, 0.9):
                grid.append(dict(alpha=a, l1_ratio=l1, max_iter=1000, fit_intercept=fi)) # ElasticNet
    grid.append(dict(alpha=1e-3, l1_ratio=0.5, max_iter=1000, fit_intercept=True, tol=1e-10))
    grid.append(dict(alpha=1e-2, l1_ratio=0.5, max_iter=3, fit_intercept=True))           # max_iter exhausted
    coefs, icpts, r2s, mses, n_iters = [], [], [], [], []
    for kw in grid:
        g = sglm.GLM("Gaussian", **dict(kw))
        g.fit(Xd, y)
        coefs.append(np.asarray(g.coef_, dtype=np.float64))
        icpts.append(float(g.intercept_))
        r2s.append(float(g.r2_score(Xd, y)))
        mses.append(float(g.neg_mse_score(Xd, y)))
        n_iters.append(int(getattr(g.model, "n_iter_", -1) or -1))
    np.savez_compressed(os.path.join(OUT, "fits_ref.npz"), X0=X0, shifts=np.asarray(shifts),
                        keep=keep, y=y, coefs=np.stack(coefs), intercepts=np.asarray(icpts),
                        r2=np.asarray(r2s), neg_mse=np.asarray(mses), n_iter=np.asarray(n_iters))
    with open(os.path.join(OUT, "fits_ref.json"), "w") as f:
        json.dump(dict(versions=VERSIONS, grid=grid), f, indent=1)
    print("fits:", len(grid))

    # Poisson: the reference wrapper fits and then raises AttributeError
    # (backend/sglm.py:246-250), so the fixture is the estimator it constructs
    # (TweedieRegressor(power=1), sglm.py:112-115) driven to its optimum.
    X0p, shp, keepp, Xdp, yp = small_problem(seed=5, poisson=True)
    pg = [dict(alpha=a, fit_intercept=fi) for fi in (True, False) for a in (1e-4, 1e-2, 1.0)]
    pc, pi, pd2, pdef_c, pdef_i = [], [], [], [], []
    for kw in pg:
        m = TweedieRegressor(power=1, solver="newton-cholesky", tol=1e-12, max_iter=1000, **kw).fit(Xdp, yp)
        pc.append(m.coef_.copy()); pi.append(float(m.intercept_)); pd2.append(float(m.score(Xdp, yp)))
        m2 = TweedieRegressor(power=1, **kw).fit(Xdp, yp)   # reference defaults: lbfgs, tol 1e-4
        pdef_c.append(m2.coef_.copy()); pdef_i.append(float(m2.intercept_))
    raised = ""
    try:
        quiet(sglm.GLM("Poisson", alpha=1e-2).fit, Xdp, yp)
    except Exception as e:  # documents the reference behaviour
        raised = type(e).__name__
    np.savez_compressed(os.path.join(OUT, "poisson_ref.npz"), X0=X0p, shifts=np.asarray(shp), keep=keepp,
                        y=yp, coefs=np.stack(pc), intercepts=np.asarray(pi), d2=np.asarray(pd2),
                        coefs_default=np.stack(pdef_c), intercepts_default=np.asarray(pdef_i))
    with open(os.path.join(OUT, "poisson_ref.json"), "w") as f:
        json.dump(dict(versions=VERSIONS, grid=pg, reference_wrapper_raises=raised), f, indent=1)
    print("poisson fits:", len(pg), "reference wrapper raises:", raised)


# --------------------------------------------------------------------------- #
# 3. CV grid through the reference cv_glm_mult_params
# --------------------------------------------------------------------------- #
def save_cv():
    X0, shifts, keep, Xd, y = small_problem(seed=11, T=4000, P=4, h=3)
    cv_idx = orc.synth_folds(Xd.shape[0], 4, seed=11, group=200)
    blob = dict(X0=X0, shifts=np.asarray(shifts), keep=keep, y=y)
    for k, (a, b) in enumerate(cv_idx):
        blob[f"train{k}"] = a
        blob[f"test{k}"] = b
    meta = dict(versions=VERSIONS, n_folds=len(cv_idx), runs=[])
    for tag, score_method, lists, fixed in [
        ("enet_mse", "mse", dict(alpha=[1e-3, 1e-2, 1e-1], l1_ratio=[0.1, 0.5, 1.0]),
         dict(max_iter=1000, fit_intercept=True)),
        ("mixed_r2", "r2", dict(alpha=[0, 1e-2, 10.0], l1_ratio=[0, 0.5]),
         dict(max_iter=1000, fit_intercept=True)),
        ("noint_roll", "r2", dict(alpha=[1e-2, 1.0], l1_ratio=[0, 0.3], roll=[0, 5]),
         dict(max_iter=1000, fit_intercept=False)),
    ]:
        kw_lst = sglm_cv.generate_mult_params(lists, fixed)
        kw_snapshot = [dict(k) for k in kw_lst]
        res = quiet(sglm_cv.cv_glm_mult_params, Xd, y, cv_idx, "Gaussian", kw_lst,
                    score_method=score_method)
        full = res["full_cv_results"]
        for j, r in enumerate(full):
            blob[f"{tag}_coefs{j}"] = r["cv_coefs"]
            blob[f"{tag}_icpt{j}"] = r["cv_intercepts"]
            blob[f"{tag}_tr{j}"] = r["cv_scores_train"]
            blob[f"{tag}_te{j}"] = r["cv_scores_test"]
            blob[f"{tag}_agg{j}"] = np.asarray([r["cv_mean_score_train"], r["cv_mean_score"],
                                                r["cv_std_score"], r["cv_R2_score"], r["cv_mse_score"]])
            blob[f"{tag}_fullcoef{j}"] = np.asarray(r["model"].coef_)
            blob[f"{tag}_fullicpt{j}"] = np.asarray(float(r["model"].intercept_))
        meta["runs"].append(dict(tag=tag, score_method=score_method, kwargs=kw_snapshot,
                                 best_params=res["best_params"], best_score=float(res["best_score"]),
                                 best_score_std=float(res["best_score_std"]),
                                 result_kwargs=[r["glm_kwargs"] for r in full]))
        print(tag, "sets:", len(kw_lst), "best:", res["best_params"], res["best_score"])
    np.savez_compressed(os.path.join(OUT, "cv_ref.npz"), **blob)
    with open(os.path.join(OUT, "cv_ref.json"), "w") as f:
        json.dump(meta, f, indent=1)


# --------------------------------------------------------------------------- #
# 4. second-generation lag builder (sglm/sglm/features/setup_model_fit.py:43-96)
# --------------------------------------------------------------------------- #
def save_by_dict():
    """The module needs the whole `sglm` package (and lab-specific imports) to import, so the two
    functions are compiled UNMODIFIED from the reference file's AST at run time."""
    import ast
    path = "/root/reference/sglm/sglm/features/setup_model_fit.py"
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("timeshift_vals_by_dict", "X_cols_dict_to_default")]
    ns = {"np": np, "pd": pd}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    rng = np.random.default_rng(21)
    df = pd.DataFrame({"a": rng.standard_normal(60), "b": (rng.random(60) < 0.2).astype(float),
                       "c": np.arange(60.0), "lbl": rng.integers(0, 3, 60)})
    cases = [dict(d={"a": (-2, 3), "b": (-1, 1)}, keep_nans=False), dict(d={"b": (-3, 0), "a": (0, 2)}, keep_nans=True),
             dict(d={"c": (-1, 4)}, keep_nans=False)]
    blob, meta = {"df": df.to_numpy(dtype=np.float64)}, dict(versions=VERSIONS, columns=list(df.columns), cases=[])
    for i, c in enumerate(cases):
        out, names = ns["timeshift_vals_by_dict"](df, dict(c["d"]), keep_nans=c["keep_nans"])
        blob[f"out{i}"] = out.to_numpy(dtype=np.float64)
        blob[f"idx{i}"] = np.asarray(out.index)
        meta["cases"].append(dict(d={k: list(v) for k, v in c["d"].items()}, keep_nans=c["keep_nans"],
                                  names=names, out_columns=list(out.columns)))
    meta["default"] = {k: list(v) for k, v in ns["X_cols_dict_to_default"]({"a": (0, 0), "b": None, "c": (-1, 2)}).items()}
    np.savez_compressed(os.path.join(OUT, "by_dict_ref.npz"), **blob)
    with open(os.path.join(OUT, "by_dict_ref.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("by_dict cases:", len(cases))


def save_preprocess():
    """zscore / diff / detrend_data of the reference's preprocessing module (backend/sglm_pp.py:105-190, :522-545) on
    seeded inputs: ungrouped and grouped (interleaved group rows) rolling min-max, NaNs inside the series."""
    rng = np.random.default_rng(33)
    n = 900
    x = np.cumsum(rng.standard_normal(n)) + 0.01 * np.arange(n)
    x[[100, 101, 560]] = np.nan
    grp = rng.integers(0, 3, n)                     # interleaved groups
    grp2 = (np.arange(n) // 300)                    # contiguous groups
    df = pd.DataFrame({"sig": x, "g": grp, "h": grp2, "other": rng.standard_normal(n)})
    blob = {"sig": x, "g": grp, "h": grp2, "other": df["other"].to_numpy()}
    meta = dict(versions=VERSIONS, cases=[])
    for i, (cols, window) in enumerate([([], 20), (["g"], 10), (["h"], 25), (["g", "h"], 6)]):
        out = sglm_pp.detrend_data(df, "sig", cols, window)
        blob[f"detrend{i}"] = out.to_numpy(dtype=np.float64)
        idx = out.index.to_frame(index=False).to_numpy() if isinstance(out.index, pd.MultiIndex) else np.asarray(out.index)
        blob[f"detrend_idx{i}"] = np.asarray(idx, dtype=np.int64)
        meta["cases"].append(dict(grouping_cols=cols, window=window, index_names=list(out.index.names)))
    Z = rng.standard_normal((200, 5)) * np.array([1, 10, 0.1, 3, 7]) + np.array([0, 5, -2, 1, 100])
    blob["Z"] = Z
    blob["zscore_np"] = sglm_pp.zscore(Z)
    blob["zscore_df"] = sglm_pp.zscore(pd.DataFrame(Z)).to_numpy()
    blob["diff1"] = sglm_pp.diff(Z)
    blob["diff2_cols"] = sglm_pp.diff(Z, diff_inx=[1, 3], n=2)
    blob["diff1_append"] = sglm_pp.diff(Z, diff_inx=[0, 4], append_to_base=True)
    np.savez_compressed(os.path.join(OUT, "preprocess_ref.npz"), **blob)
    with open(os.path.join(OUT, "preprocess_ref.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("preprocess cases:", len(meta["cases"]))


if __name__ == "__main__":
    if "--only-preprocess" in sys.argv:
        save_preprocess()
        sys.exit(0)
    if "--only-by-dict" in sys.argv:
        save_by_dict()
        sys.exit(0)
    save_by_dict()
    save_preprocess()
    save_gather()
    save_fits()
    save_cv()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
