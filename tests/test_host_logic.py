"""CPU tests of the host-side logic: C-ABI surface, column maps, GLM dispatch, parameter
grids, CPU passthrough helpers, and the rule that the product never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from conftest import PKG, ROOT, load_golden
from oracle import sglm_oracle as orc

import _sglm_native as nat
import sglm
import sglm_
import sglm_cv
import sglm_dist
import sglm_ez
import sglm_pp


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sglm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sglm_[a-z0-9_]+)\s*\(", hdr)))
    lib = ctypes.CDLL(os.path.join(PKG, "libsglm_b200.so"))
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(nat.exported_symbols()) == declared          # binding covers exactly the header
    lib.sglm_version.restype = ctypes.c_int
    assert lib.sglm_version() >= 100
    assert nat.lib() is not None


def test_abi_argument_validation_without_gpu():
    """Shape / null checks are evaluated before any CUDA call, so they are testable on CPU."""
    lib = nat.lib()
    rc = lib.sglm_timeshift_f64_ranged(None, -1, 3, 3, None, None, 2, 0, 1, 0, None, 2, None)
    assert rc == -2 and b"bad shape" in lib.sglm_last_error()
    rc = lib.sglm_timeshift_f64_ranged(None, 10, 3, 3, None, None, 2, 0, 1, 0, None, 2, None)
    assert rc == -1                                             # null pointers
    assert lib.sglm_suffstats_workspace_bytes(1000, 10, 1, 0, None) == 0        # n_sets out of range
    assert lib.sglm_suffstats_workspace_bytes(1000, 10, 1, 2, None) > 0
    assert lib.sglm_ridge_workspace_bytes(10, 16, 3) == 3 * 11 * 16 * 8
    rc = lib.sglm_quadform_f64(None, 4, 8, None, 8, 1, None, None)
    assert rc == -2                                             # lda < n


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nat.SglmNativeError):
        sglm_pp.timeshift(np.zeros((4, 2)))
    with pytest.raises(nat.SglmNativeError):
        sglm.GLM("Gaussian", alpha=0.1, l1_ratio=0.5).fit(np.zeros((4, 2)), np.zeros(4))


def test_product_never_imports_the_oracle():
    for fn in os.listdir(PKG):
        if fn.endswith(".py"):
            src = open(os.path.join(PKG, fn)).read()
            assert "oracle" not in src.replace("the CPU oracle", ""), fn
    for fn in os.listdir(os.path.join(PKG, "csrc")):
        assert "oracle" not in open(os.path.join(PKG, "csrc", fn)).read(), fn


def test_column_map_matches_oracle_layout_contract():
    for n_cols, inx, shifts, keep in [(4, [], [-1, 0, 1], True), (4, [0, 3], [-1, 0, 1], True),
                                      (6, [2], [2, 0, -2], False), (6, [1, 4], [0, -3, -2, -1, 1, 2, 3], True),
                                      (6, [0, 5], [-1, 1, 0], True), (3, [1], [1, 1, 0], True)]:
        src, sh, sizes = sglm_pp.build_column_map(n_cols, inx, shifts, keep)
        osrc, osh = orc.column_map(n_cols, inx, shifts, keep)
        assert np.array_equal(src, osrc) and np.array_equal(sh, osh)
        assert sum(sizes) == len(src)
    with pytest.raises(IndexError):
        sglm_pp.build_column_map(3, [3], [1])


def test_generate_mult_params_matches_reference_fixture():
    _, meta = load_golden("cv_ref")
    run = meta["runs"][0]
    got = sglm_cv.generate_mult_params(dict(alpha=[1e-3, 1e-2, 1e-1], l1_ratio=[0.1, 0.5, 1.0]),
                                       dict(max_iter=1000, fit_intercept=True))
    assert got == run["kwargs"]                                 # order: fixed keys first, last key fastest
    assert [list(d) for d in got] == [list(d) for d in run["kwargs"]]
    assert sglm_cv.generate_mult_params({"alpha": reversed([1, 2])}) == [{"alpha": 2}, {"alpha": 1}]
    assert got == orc.generate_mult_params(dict(alpha=[1e-3, 1e-2, 1e-1], l1_ratio=[0.1, 0.5, 1.0]),
                                           dict(max_iter=1000, fit_intercept=True))


def test_glm_dispatch_mirrors_reference_table():
    """backend/sglm.py:95-126 — estimator chosen from (model_name, alpha, l1_ratio)."""
    G = sglm.GLM
    assert type(G("Gaussian", alpha=0, l1_ratio=0.5, max_iter=10).model) is sglm.LinearRegression
    assert type(G("Normal", alpha=1.0, l1_ratio=0, max_iter=10).model) is sglm.Ridge
    assert type(G("Gaussian", alpha=1.0, l1_ratio=1).model) is sglm.Lasso
    assert type(G("Gaussian", alpha=1.0, l1_ratio=0.3).model) is sglm.ElasticNet
    assert type(G("Gaussian").model) is sglm.ElasticNet and G("Gaussian").model.alpha == 1.0
    assert type(G("PCA Normal").model) is sglm.LinearRegression
    p = G("Poisson", alpha=0.5)
    assert type(p.model) is sglm.TweedieRegressor and p.model.power == 1 and p.kwargs["power"] == 1
    assert G("Gamma").model.power == 2
    assert G("Gaussian", alpha=1.0, l1_ratio=0, max_iter=7).kwargs == {"alpha": 1.0, "max_iter": 7}
    with pytest.raises(KeyError):
        G("Gaussian", alpha=0)                                   # l1_ratio / max_iter missing (:98-99)
    with pytest.raises(TypeError):
        G("Gaussian", reg_lambda=0.1)                            # stale pyglmnet kwarg (test_sglm.py:25)
    with pytest.raises(NameError):
        G("Nope")                                                # undefined NotYetImplementedError (:126)
    with pytest.raises(NotImplementedError):
        G("Logistic")
    g = G("Gaussian", beta0_=1.5, beta_=np.arange(3.0), alpha=0.1, l1_ratio=0.5)
    assert g.kwargs["warm_start"] is True and g.model.warm_start is True
    assert np.array_equal(g.model.coef_, np.arange(3.0)) and g.beta0_ == 1.5
    assert G("Gaussian", score_method="r2").score.__func__ is G.r2_score
    assert G("Gaussian", score_method="anything").score.__func__ is G.neg_mse_score
    assert sglm_.GLM is sglm.GLM and sglm_.calc_R2 is sglm.calc_R2
    assert sglm.calc_R2(np.array([1.0, -1.0]), np.array([2.0, -2.0])) == 0.75
    assert sglm.calc_R2(np.array([1.0]), np.array([0.0])) == 0


def test_cpu_passthrough_helpers():
    X = np.array([[0, -1, 0], [1, 1, 0], [0, 1, 0], [2, 3, 4]])
    assert np.all(sglm_pp.diff(X) == np.array([[1, 2, 0], [-1, 0, 0], [2, 2, 4]]))       # test_sglm_pp.py:163
    with np.errstate(all="ignore"):
        z = sglm_pp.zscore(X.astype(float)[:, :2])
    assert np.allclose(z.mean(0), 0) and np.allclose(z.std(0), 1)
    df = pd.DataFrame(X, columns=list("ABC"))
    d = sglm_pp.diff(df)
    assert list(d.columns) == ["A_diff", "B_diff", "C_diff"] and list(d.index) == [1, 2, 3]
    d2 = sglm_pp.diff(df, [0], append_to_base=True)
    assert list(d2.columns) == ["A", "B", "C", "A_diff"] and np.isnan(d2["A_diff"].iloc[0])
    assert sglm_pp.get_column_nums(df, ["C", "A"]) == [2, 0]
    ids = sglm_pp.bucket_ids_by_timeframe(100, timesteps_per_bucket=20)
    assert ids.max() == 19                                       # reference quirk: divides by num_buckets (=5)
    np.random.seed(1)
    cv = sglm_pp.cv_idx_from_bucket_ids(ids, np.zeros((100, 1)), num_folds=3, test_size=0.25)
    assert len(cv) == 3 and all(len(set(a) & set(b)) == 0 for a, b in cv)
    assert sglm_ez.add_timeshifts_to_col_list(["A", "B"], ["A"], -1, 2) == ["A", "B", "A_-1", "A_1", "A_2"]
    assert sglm_ez._shift_list(-2, 1) == [0, -2, -1, 1]


def test_shard_indices_partition():
    cost = [sglm_dist.default_cost(dict(alpha=a, l1_ratio=l)) for a in (1e-3, 1e-1, 0.0) for l in (0.0, 0.5, 1.0)]
    parts = [sglm_dist.shard_indices(len(cost), 4, r, cost) for r in range(4)]
    assert sorted(i for p in parts for i in p) == list(range(len(cost)))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    heavy = int(np.argmax(cost))
    assert heavy in parts[0]                                     # most expensive set dealt first


def test_cd_launch_plan_parsing():
    """Launch plan of the coordinate-descent grid (host logic only; the library is loaded to ask which
    cluster shapes are compiled in)."""
    import _engine as eng
    old = (eng.CD_PLAN, eng.CD_GROUP, eng.CD_CLUSTER)
    try:
        eng.CD_GROUP = eng.CD_CLUSTER = None
        eng.CD_PLAN = "4x2@0.3,0x0"
        assert eng._cd_plan(2000, 1500) == [(0, 450, 4, 2), (450, 1500, 0, 0)]
        eng.CD_PLAN = "2x8@0.1,4x4@0.2,1x1"
        assert eng._cd_plan(2000, 1000) == [(0, 100, 2, 8), (100, 300, 4, 4), (300, 1000, 1, 1)]
        eng.CD_PLAN = "4x8"
        assert eng._cd_plan(96, 10) == [(0, 10, 4, 1)]           # 3 coordinate blocks cannot feed 8 CTAs
        eng.CD_PLAN = None
        assert eng._cd_plan(410, 1500) == [(0, 1500, 0, 0)]       # narrow design: one CTA per model
        assert eng._cd_plan(410, 1) == [(0, 1, 1, 4)]             # a single fit: one 4-CTA cluster
        assert eng._cd_plan(100, 1) == [(0, 1, 0, 0)]
        assert eng._cd_plan(2000, 1500)[0][2:] == (4, 4)
        # a whole grid: the heaviest group of every problem on (4,4), the rest of the heavy 30 % on (4,2)
        assert eng._cd_plan(2000, 1500, 6) == [(0, 24, 4, 4), (24, 450, 4, 2), (450, 1500, 0, 0)]
        # shares of a grid (one of several GPUs): the chain-critical head on the widest shape, then what the SMs allow
        assert eng._cd_plan(2000, 750, 6) == [(0, 24, 2, 4), (24, 72, 4, 4), (72, 225, 4, 2), (225, 750, 0, 0)]
        assert eng._cd_plan(2000, 375, 6) == [(0, 24, 2, 4), (24, 112, 4, 4), (112, 375, 0, 0)]
        assert eng._cd_plan(2000, 188, 6) == [(0, 6, 1, 8), (6, 56, 2, 4), (56, 188, 0, 0)]
        eng.CD_PLAN = "4x4#24,4x2#100,0x0"
        assert eng._cd_plan(2000, 1500) == [(0, 24, 4, 4), (24, 124, 4, 2), (124, 1500, 0, 0)]
        eng.CD_PLAN = None
        eng.CD_PLAN = "4x2"
        assert eng._cd_plan(6000, 100) == [(0, 100, 4, 4)]       # slice of 3000 columns x 4 models does not fit: wider cluster
        assert eng._cd_plan(20000, 100)[0][2] in (1, 2)          # ... then smaller groups
        eng.CD_PLAN = "3x2"
        with pytest.raises(Exception):
            eng._cd_plan(2000, 10)
    finally:
        eng.CD_PLAN, eng.CD_GROUP, eng.CD_CLUSTER = old


def test_union_of_overlapping_call_intervals():
    """bench.py times the coordinate-descent grid as the union of its concurrent launches."""
    old = nat.last_intervals
    try:
        nat.last_intervals = [("a", 0.0, 10.0), ("b", 5.0, 12.0), ("a", 20.0, 21.0), ("c", 0.0, 100.0)]
        assert nat.union_ms(("a", "b")) == pytest.approx(13.0)
        assert nat.union_ms(("a",)) == pytest.approx(11.0)
        assert nat.union_ms(("zzz",)) == 0.0
    finally:
        nat.last_intervals = old


def test_exporters_and_gen2_layout_without_gpu(tmp_path):
    """Result formats (sglm_save.py:7-68; er_refactored_from_scratch_cleanup.py:528-537) and the second-generation
    package tree need no GPU: pickles of the result container round-trip, file names follow the drivers' rule."""
    import pickle
    import sglm_save
    store = sglm_save.GLM_data(str(tmp_path), "fits.pkl")
    store.set_uid("u"); store.set_filename("f"); store.set_basedata("b"); store.set_X_cols(["a"]); store.set_timeshifts(-3, 4)
    store.set_gss_info(5, 0.2, 0.25, gssid=3)
    store.append_fit_results("resp", {"alpha": 0.1}, glm_model=None, scores={"tr_witi": 0.5}, dropped_cols=["x"], gssids=[1])
    store.save()
    store.save()                                     # second save refuses to overwrite, as the reference does
    with open(tmp_path / "fits.pkl", "rb") as f:
        back = pickle.load(f)
    assert back.data["negorder"] == -3 and back.data["posorder"] == 4 and back.data["gss_info"]["gssid"] == 3
    fr = back.data["fit_results"][0]
    assert fr["scores"]["tr_witi"] == 0.5 and fr["scores"]["holdout_noiti"] is None and fr["gss_mse"] is None
    other = sglm_save.GLM_data(str(tmp_path), "fits.pkl")
    other.load()
    assert other.data.data["uid"] == "u"             # the reference's load() assigns the unpickled OBJECT to .data
    assert sglm_save.model_file_stem("r1", {"alpha": 0.5, "l1_ratio": 0.1, "max_iter": 1000}) == "r1_alpha_0.5_l1_ratio_0.1_max_iter_1000"
    tree = os.path.join(PKG, "gen2", "sglm")
    for rel in ("models/sglm.py", "models/sglm_cv.py", "models/split_data.py", "models/eval.py", "models/train_model.py",
                "features/sglm_pp.py", "features/setup_model_fit.py", "data/save_results.py"):
        assert os.path.exists(os.path.join(tree, rel)), rel


def test_lag_recipe_window_and_offsets():
    """Host logic of the never-built lag design (_engine.LagRecipe): the window of base rows every column reads, the per-column
    offsets (design row t of column c = window row t + off[c]) and the cases that must fall back to building the design."""
    import _engine as eng

    class Base:                      # only .shape is consulted by window()
        def __init__(self, T, P):
            self.shape = (T, P)

    T, P = 1000, 3
    sh = np.array([0, -2, -1, 1, 3, 3], dtype=np.int32)          # out[t, c] = base[t - sh[c], src[c]]
    src = np.array([0, 1, 2, 0, 1, 2], dtype=np.int32)
    # rows kept by dropna: t - sh >= 0 and t - sh < T for every column -> [3, 998)
    r = eng.LagRecipe(Base(T, P), src, sh, 3, 998, float("nan"))
    assert r.shape == (995, 6)
    u0, n_u, off = r.window()
    assert u0 == 0 and n_u == 1000
    # design row 0 is base row 3 - sh[c]
    assert list(off) == [3, 5, 4, 2, 0, 0]
    assert off.min() >= 0 and off.max() + r.shape[0] <= n_u
    # an interior slice of rows (what one of several GPUs holds): the window moves with it
    r2 = eng.LagRecipe(Base(T, P), src, sh, 400, 650, float("nan"))
    u0, n_u, off = r2.window()
    assert (u0, n_u) == (397, 255) and list(off) == [3, 5, 4, 2, 0, 0]
    # columns that read outside the base signals (fill values would be needed): no window -> the design is built
    assert eng.LagRecipe(Base(T, P), src, sh, 0, T, 0.0).window() is None
    assert eng.LagRecipe(Base(T, P), src, sh, 2, 998, 0.0).window() is None
    assert eng.LagRecipe(Base(T, P), src, sh, 3, 999, 0.0).window() is None
    # empty designs
    assert eng.LagRecipe(Base(T, P), src, sh, 5, 5, 0.0).window() is None
    assert eng.LagRecipe(Base(T, P), src[:0], sh[:0], 0, T, 0.0).window() is None
