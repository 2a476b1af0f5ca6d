"""GPU parity tests: the CUDA path (through the Python mirror and the C ABI of
libsglm_b200.so) against the oracle and the golden fixtures produced from the unmodified
reference.  Tolerances: design matrices bit-exact; coefficients rel 1e-4
(||dw||_inf / ||w||_inf) at identical tol/max_iter, direct solvers 1e-7; scores abs 1e-6."""
import numpy as np
import pandas as pd
import pytest

from conftest import coef_rel_err, load_golden
from oracle import sglm_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import _engine as eng  # noqa: E402
import _sglm_native as nat  # noqa: E402
import sglm  # noqa: E402
import sglm_cv  # noqa: E402
import sglm_ez  # noqa: E402
import sglm_pp  # noqa: E402


def _kw(kwargs):
    kw = dict(kwargs)
    if kw.get("fill_value") == "nan":
        kw["fill_value"] = np.nan
    return kw


def _design(blob):
    Xd = orc.timeshift_multiple(blob["X0"], shift_amt_list=[int(s) for s in blob["shifts"]])
    return Xd[blob["keep"]]


# ------------------------------------------------------------------ gather
def test_gather_golden_bit_exact():
    blob, meta = load_golden("gather_ref")
    for i, case in enumerate(meta["cases"]):
        X, want = blob[f"x{i}"], blob[f"r{i}"]
        kw = _kw(case["kwargs"])
        got = sglm_pp.timeshift_multiple(X, **kw) if "shift_amt_list" in kw else sglm_pp.timeshift(X, **kw)
        assert got.shape == want.shape, case["name"]
        assert got.dtype == want.dtype, case["name"]
        assert got.tobytes() == want.tobytes(), case["name"]


def test_gather_reference_unit_tests_dataframe():
    """backend/test/test_sglm_pp.py:115-151 re-stated against the drop-in module."""
    _, meta = load_golden("gather_ref")
    df = pd.DataFrame(np.arange(20).reshape(5, 4), columns=list("ABCD"))
    a = np.arange(20).reshape(5, 4)
    r1 = sglm_pp.timeshift_multiple(df, shift_amt_list=[-1, 0, 1], unshifted_keep_all=True, fill_value=0)
    assert list(r1.columns) == meta["df_names_all"]
    assert np.all(r1.values == orc.timeshift_multiple(a, shift_amt_list=[-1, 0, 1], fill_value=0))
    inx = sglm_pp.get_column_nums(df, ["A", "D"])
    r2 = sglm_pp.timeshift_multiple(df, shift_inx=inx, shift_amt_list=[-1, 0, 1], fill_value=0)
    assert list(r2.columns) == meta["df_names_sub"]
    assert np.all(r2.values == orc.timeshift_multiple(a, shift_inx=[0, 3], shift_amt_list=[-1, 0, 1], fill_value=0))
    r3 = sglm_pp.timeshift(df, shift_inx=inx, shift_amt=1, fill_value=0, keep_non_inx=True)
    assert list(r3.columns) == list("ABCD")
    assert np.all(r3.values == orc.timeshift(a, shift_inx=[0, 3], shift_amt=1, fill_value=0, keep_non_inx=True))
    assert np.all(sglm_pp.timeshift(df, shift_amt=0) == df)
    with pytest.raises(ValueError):
        sglm_pp.get_column_nums(pd.DataFrame(np.zeros((2, 3)), columns=["A", "B", "A"]), ["A"])


@pytest.mark.parametrize("T,P,h,inx", [(5000, 7, 12, [0, 2, 5]), (4097, 10, 20, []), (300, 3, 1, [1]),
                                       (2048, 33, 5, list(range(0, 33, 2)))])
def test_gather_random_vs_oracle_bit_exact(T, P, h, inx):
    rng = np.random.default_rng(T + P)
    X = rng.standard_normal((T, P))
    X[rng.integers(0, T, 20), rng.integers(0, P, 20)] = np.nan
    shifts = [0] + list(range(-h, 0)) + list(range(1, h + 1))
    want = orc.timeshift_multiple(X, shift_inx=inx, shift_amt_list=shifts)
    got = sglm_pp.timeshift_multiple(X, shift_inx=inx, shift_amt_list=shifts)
    assert got.tobytes() == want.tobytes()
    got_t = sglm_pp.timeshift_multiple(torch.from_numpy(X).cuda(), shift_inx=inx, shift_amt_list=shifts)
    assert got_t.is_cuda and got_t.cpu().numpy().tobytes() == want.tobytes()


def test_gather_nan_payload_and_negative_zero_preserved():
    X = np.zeros((64, 2))
    bits = X.view(np.uint64)
    bits[5, 0] = 0x7FF8DEADBEEF0001      # NaN with a payload
    bits[6, 1] = 0x8000000000000000      # -0.0
    got = sglm_pp.timeshift(X, shift_amt=3)
    want = orc.timeshift(X, shift_amt=3)
    assert got.tobytes() == want.tobytes()
    assert got.view(np.uint64)[8, 0] == 0x7FF8DEADBEEF0001
    assert got.view(np.uint64)[0, 0] == nat.NAN_BITS


def test_gather_direct_kernel_paths():
    """Very wide source (window does not fit shared memory) and |shift| >= T use the direct kernel."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((257, 3000))
    shifts = [0, -40, 40, 7]
    want = orc.timeshift_multiple(X, shift_inx=[0, 1500, 2999], shift_amt_list=shifts)
    got = sglm_pp.timeshift_multiple(X, shift_inx=[0, 1500, 2999], shift_amt_list=shifts)
    assert got.tobytes() == want.tobytes()
    Y = rng.standard_normal((10, 4))
    for a in (10, -10, 25, 9, -9):
        assert sglm_pp.timeshift(Y, shift_amt=a).tobytes() == orc.timeshift(Y, shift_amt=a).tobytes()


def test_gather_generic_abi_entry_and_strided_source():
    """sglm_timeshift_f64 (range recovered on device) + a row-strided source view."""
    rng = np.random.default_rng(9)
    big = torch.from_numpy(rng.standard_normal((500, 12))).cuda()
    view = big[:, 2:9]                               # ldx = 12, 7 columns
    src = np.array([0, 3, 6, 1], dtype=np.int32)
    sh = np.array([0, 2, -3, 5], dtype=np.int32)
    d = torch.from_numpy(np.concatenate([src, sh])).cuda()
    out = torch.empty((500, 4), dtype=torch.float64, device="cuda")
    nat.call("sglm_timeshift_f64", nat.ptr(view), 500, 7, 12, nat.ptr(d[:4]), nat.ptr(d[4:]), 4,
             nat.NAN_BITS, nat.ptr(out), 4, nat.stream_ptr())
    want = orc.timeshift_c(view.cpu().numpy(), src, sh)
    assert out.cpu().numpy().tobytes() == want.tobytes()


def test_gather_roundtrip_property_full_size_rows():
    """shift by +k then by -k restores the interior rows (size-independent property)."""
    T, P, k = 200_000, 8, 17
    X = torch.randn((T, P), dtype=torch.float64, device="cuda")
    back = sglm_pp.shift(sglm_pp.shift(X, k), -k)
    assert torch.equal(back[: T - k], X[: T - k])
    assert bool(torch.isnan(back[T - k:]).all())


def test_concat_crop_helpers():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((9, 3))
    blanks = rng.standard_normal((2, 3))
    assert np.array_equal(sglm_pp.concat_start_crop_end(blanks, X), np.concatenate([blanks, X])[:-2])
    assert np.array_equal(sglm_pp.concat_end_crop_start(blanks, X), np.concatenate([X, blanks])[2:])


# ------------------------------------------------------------------ statistics
@pytest.mark.parametrize("T,C,n_y,weighted,ld_pad", [(3000, 45, 1, False, 0), (2500, 130, 2, True, 0),
                                                      (1000, 37, 1, True, 3), (70, 300, 1, False, 1),
                                                      (5, 4, 1, False, 0)])
def test_suffstats_vs_numpy(T, C, n_y, weighted, ld_pad):
    rng = np.random.default_rng(T * 7 + C)
    Xfull = rng.standard_normal((T, C + ld_pad)) + 0.3
    X = Xfull[:, :C]
    Y = rng.standard_normal((T, n_y))
    Xd = torch.from_numpy(Xfull).cuda()[:, :C]
    Yd = torch.from_numpy(Y).cuda()
    Z = np.concatenate([X, Y, np.ones((T, 1))], axis=1)
    if weighted:
        W = np.stack([np.ones(T), (rng.random(T) < 0.2).astype(float), rng.integers(0, 3, T).astype(float)])
        W[1, : T // 2] = 0.0                         # long all-zero stretch -> skipped tiles
        G = eng.suffstats(Xd, Yd, torch.from_numpy(W).cuda(), [T, W[1].sum(), W[2].sum()])
        want = np.stack([(Z * w[:, None]).T @ Z for w in W])
    else:
        G = eng.suffstats(Xd, Yd)
        want = (Z.T @ Z)[None]
    n_aug = C + n_y + 1
    got = G[:, :, :n_aug].cpu().numpy()
    scale = np.abs(want).max()
    assert np.max(np.abs(got - want)) <= 1e-12 * scale
    assert np.array_equal(got, np.transpose(got, (0, 2, 1)))          # exactly symmetric


def _mixed_design(T, C, seed):
    """Photometry-like columns: 0/1 indicators, small integers, dyadic fractions, general reals."""
    rng = np.random.default_rng(seed)
    X = np.empty((T, C))
    for c in range(C):
        kind = c % 5
        if kind in (0, 1):
            X[:, c] = (rng.random(T) < 0.03)
        elif kind == 2:
            X[:, c] = rng.integers(-5, 6, T)
        elif kind == 3:
            X[:, c] = rng.integers(0, 1024, T) / 1024.0
        else:
            X[:, c] = rng.standard_normal(T) * 10.0 ** rng.integers(-3, 4)
    return X


@pytest.mark.parametrize("T,C,check", [(1000, 40, True), (5000, 300, True), (40_000, 530, False)])
def test_tensor_core_gram_matches_fp64(T, C, check):
    """tcgen05 int8 digit-plane Gram == fp64 Gram (numpy / DMMA) to fp64 accuracy, for the full
    set and two row subsets; with `check`, the tcgen05 GEMM is also compared bit-for-bit with
    the CUDA-core integer GEMM run on the same digit planes."""
    X = _mixed_design(T, C, T + C)
    rng = np.random.default_rng(1)
    y = rng.standard_normal((T, 2))
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    sub1 = np.sort(rng.choice(T, T // 5, replace=False))
    sub2 = np.arange(T // 3, T // 3 + 257)
    sets = [None, torch.from_numpy(sub1).cuda(), torch.from_numpy(sub2).cuda()]
    G, colS = eng.suffstats_tc(Xd, Yd, sets)
    n_aug = C + 3
    assert colS[0] == 1 and colS[2] == 1 and colS[4] == 8 and colS[-1] == 1      # indicators / ints need one plane
    assert colS[3] <= 2
    Z = np.concatenate([X, y, np.ones((T, 1))], axis=1)
    for s, rows in enumerate([np.arange(T), sub1, sub2]):
        want = Z[rows].T @ Z[rows]
        got = G[s, :, :n_aug].cpu().numpy()
        scale = np.sqrt(np.outer(np.diag(want), np.diag(want))) + 1e-300
        assert np.max(np.abs(got - want) / scale) < 1e-13, s
        assert np.array_equal(got, got.T)
    # integer-valued columns are reproduced exactly
    ints = [c for c in range(C) if c % 5 in (0, 1, 2)]
    got0 = G[0, :, :n_aug].cpu().numpy()
    assert np.array_equal(got0[np.ix_(ints, ints)], (Z[:, ints].T @ Z[:, ints]))
    W = torch.zeros((3, T), dtype=torch.float64, device="cuda")
    W[0] = 1.0
    W[1, torch.from_numpy(sub1).cuda()] = 1.0
    W[2, torch.from_numpy(sub2).cuda()] = 1.0
    Gd = eng.suffstats(Xd, Yd, W, [T, len(sub1), len(sub2)])
    d = (G - Gd).abs().max().item() / Gd.abs().max().item()
    assert d < 1e-13
    if check:
        G2, _ = eng.suffstats_tc(Xd, Yd, sets, check_gemm=True)
        assert torch.equal(G, G2)


def test_tensor_core_gram_cell_decomposition_is_exact():
    """Overlapping row sets (full data + random, intersecting folds): the GEMM over the disjoint cells of
    the induced partition + integer cell sums must give the SAME bits as one GEMM pass per set, for the
    tcgen05 kernel and for the CUDA-core check GEMM."""
    T, C = 6000, 150
    X = _mixed_design(T, C, 5)
    rng = np.random.default_rng(8)
    y = rng.standard_normal((T, 1))
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    groups = np.arange(T) // 50
    sets = [None]
    for f in range(4):
        pick = rng.permutation(groups.max() + 1)[: (groups.max() + 1) // 4]           # random 25 % of the groups
        sets.append(torch.from_numpy(np.flatnonzero(np.isin(groups, pick))).cuda())
    assert eng._row_cells(sets, T) is not None
    old = eng.TC_CELLS
    try:
        eng.TC_CELLS = False
        G_sets, _ = eng.suffstats_tc(Xd, Yd, sets)
        assert not nat.last_tc_plan["cells"]
        eng.TC_CELLS = None
        G_cells, _ = eng.suffstats_tc(Xd, Yd, sets)
        assert nat.last_tc_plan["cells"] and nat.last_tc_plan["n_pos"] < 1.2 * T
        G_check, _ = eng.suffstats_tc(Xd, Yd, sets, check_gemm=True)
    finally:
        eng.TC_CELLS = old
    assert torch.equal(G_sets, G_cells) and torch.equal(G_cells, G_check)
    # disjoint sets (a K-fold partition without the full data): nothing to share, the per-set path is taken
    assert eng._row_cells([sets[1], torch.from_numpy(np.setdiff1d(np.arange(T), sets[1].cpu().numpy())).cuda()], T) is None
    Z = np.hstack([X, y, np.ones((T, 1))])
    for s_i, rows in enumerate(sets):
        Zs = Z if rows is None else Z[rows.cpu().numpy()]
        want = Zs.T @ Zs
        got = G_cells[s_i, :, :C + 2].cpu().numpy()
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


def test_cv_grid_kfold_partition_through_cells_matches_weighted_fp64_path():
    """A K-fold PARTITION (disjoint test folds that cover every row): the cells of the tensor-core Gram are the
    folds themselves and the full-data statistics are their sum — same CV results as the weighted fp64 DMMA path."""
    import os
    rng = np.random.default_rng(12)
    T, C = 4000, 90
    X = _mixed_design(T, C, 9)
    y = X[:, :8].sum(1) + rng.standard_normal(T)
    perm = rng.permutation(T)
    cv_idx = [(np.sort(np.setdiff1d(perm, perm[k::4])), np.sort(perm[k::4])) for k in range(4)]
    grid = [dict(alpha=a, l1_ratio=l, max_iter=1000, fit_intercept=True) for a in (0.0, 0.01, 1.0) for l in (0.0, 0.5)]
    out = {}
    for mode in ("dmma", "tc"):
        os.environ["SGLM_GRAM"] = mode
        try:
            out[mode] = sglm_cv.cv_glm_mult_params(X, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
        finally:
            os.environ.pop("SGLM_GRAM", None)
        if mode == "tc":
            assert nat.last_tc_plan["cells"] and nat.last_tc_plan["row_lists"] == 4
            assert nat.last_tc_plan["n_pos"] <= T + 4 * 128
    assert out["tc"]["best_params"] == out["dmma"]["best_params"]
    for a, b in zip(out["tc"]["full_cv_results"], out["dmma"]["full_cv_results"]):
        assert coef_rel_err(a["cv_coefs"], b["cv_coefs"]) < 1e-9
        assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], atol=1e-9)
        assert abs(a["cv_R2_score"] - b["cv_R2_score"]) < 1e-9
        assert coef_rel_err(a["model"].coef_, b["model"].coef_) < 1e-9


def test_device_side_fold_indices_equal_the_host_split():
    """cv_idx_by_timeframe / cv_idx_by_trial_id with device="cuda" return the SAME splits as the
    reference path (same draws from the global numpy RNG), as CUDA index tensors that cv_glm_* takes directly."""
    T = 5000
    X = pd.DataFrame({"a": np.arange(T, dtype=float), "nTrial": np.arange(T) // 37, "iBlock": np.arange(T) // 900})
    for maker in (lambda dev: sglm_ez.cv_idx_by_timeframe(X, timesteps_per_bucket=20, num_folds=4, test_size=0.25, device=dev),
                  lambda dev: sglm_ez.cv_idx_by_trial_id(X, trial_id_columns=["nTrial", "iBlock"], num_folds=3, test_size=0.2, device=dev)):
        np.random.seed(7)
        host = maker(None)
        after_host = np.random.get_state()[1].copy()
        np.random.seed(7)
        dev = maker("cuda")
        assert np.array_equal(after_host, np.random.get_state()[1])             # same RNG consumption
        assert len(host) == len(dev)
        for (tr_h, te_h), (tr_d, te_d) in zip(host, dev):
            assert tr_d.is_cuda and tr_d.dtype == torch.int64
            assert np.array_equal(tr_h, tr_d.cpu().numpy()) and np.array_equal(te_h, te_d.cpu().numpy())
    # ... and a CV call on them
    rng = np.random.default_rng(0)
    Xd = rng.standard_normal((T, 6))
    y = Xd[:, 0] + 0.1 * rng.standard_normal(T)
    np.random.seed(3)
    cv_dev = sglm_ez.cv_idx_by_timeframe(Xd, timesteps_per_bucket=50, num_folds=3, test_size=0.3, device="cuda")
    np.random.seed(3)
    cv_host = sglm_ez.cv_idx_by_timeframe(Xd, timesteps_per_bucket=50, num_folds=3, test_size=0.3)
    grid = [dict(alpha=0.01, l1_ratio=0.5, max_iter=1000, fit_intercept=True)]
    a = sglm_cv.cv_glm_mult_params(Xd, y, cv_dev, "Gaussian", [dict(g) for g in grid], score_method="r2")
    b = sglm_cv.cv_glm_mult_params(Xd, y, cv_host, "Gaussian", [dict(g) for g in grid], score_method="r2")
    assert np.array_equal(a["full_cv_results"][0]["cv_coefs"], b["full_cv_results"][0]["cv_coefs"])


def test_tensor_core_gram_rejects_nan():
    X = np.random.default_rng(0).standard_normal((300, 8))
    X[17, 3] = np.inf
    with pytest.raises(ValueError):
        eng.suffstats_tc(torch.from_numpy(X).cuda(), torch.zeros((300, 1), dtype=torch.float64, device="cuda"), [None])


def test_index_counts_matches_bincount():
    rng = np.random.default_rng(4)
    idx = rng.integers(-50, 1000, 5000)
    got = eng.index_counts(idx, 1000).cpu().numpy()
    want = np.bincount(np.where(idx < 0, idx + 1000, idx), minlength=1000).astype(float)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ solvers on statistics
def test_cd_kernel_matches_oracle_gram_cd_iterate_for_iterate():
    rng = np.random.default_rng(11)
    n, p = 4000, 60
    X = rng.standard_normal((n, p)) @ (np.eye(p) + 0.5 * rng.standard_normal((p, p)) / np.sqrt(p))
    X[:, 7] = 0.0                                     # zero column (Q[j,j] == 0 branch)
    y = X @ (rng.standard_normal(p) * (rng.random(p) < 0.4)) + rng.standard_normal(n)
    for alpha, l1r, tol, mi in [(0.05, 0.5, 1e-4, 1000), (0.5, 1.0, 1e-4, 1000), (1e-3, 0.1, 1e-8, 1000),
                                (0.02, 0.9, 1e-4, 2)]:
        w_o, b_o, info_o = orc.enet_fit(X, y, alpha, l1r, True, mi, tol, use_gram=True)
        est = sglm.ElasticNet(alpha=alpha, l1_ratio=l1r, tol=tol, max_iter=mi)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            est.fit(X, y)
        assert est.n_iter_ == info_o["n_iter"], (alpha, l1r, est.n_iter_, info_o)
        assert coef_rel_err(est.coef_, w_o) < 1e-9
        assert abs(est.intercept_ - b_o) < 1e-9


def test_ridge_cholesky_multi_tile_vs_numpy():
    rng = np.random.default_rng(13)
    n, p = 3000, 203                                   # > 3 tiles of 64, ragged 32-blocks
    X = rng.standard_normal((n, p)) + 1.0
    y = rng.standard_normal(n)
    for alpha in (1e-3, 1.0, 1e3):
        for fi in (True, False):
            w_o, b_o = orc.ridge_fit(X, y, alpha, fi)
            g = sglm.GLM("Gaussian", alpha=alpha, l1_ratio=0, fit_intercept=fi, max_iter=10)
            g.fit(X, y)
            assert coef_rel_err(g.coef_, w_o) < 1e-9, (alpha, fi)
            assert abs(g.intercept_ - b_o) < 1e-9


def test_fits_vs_reference_golden():
    blob, meta = load_golden("fits_ref")
    Xd, y = _design(blob), blob["y"]
    n_iter_mismatch = 0
    import warnings
    for i, kw in enumerate(meta["grid"]):
        g = sglm.GLM("Gaussian", **dict(kw))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g.fit(Xd, y)
        kind = g.model.kind
        lim = 1e-4 if kind in ("lasso", "enet") else 1e-7
        assert coef_rel_err(g.coef_, blob["coefs"][i]) < lim, (kw, coef_rel_err(g.coef_, blob["coefs"][i]))
        assert abs(g.intercept_ - blob["intercepts"][i]) < 1e-6 * max(1.0, abs(blob["intercepts"][i])), kw
        assert abs(g.r2_score(Xd, y) - blob["r2"][i]) < 1e-6, kw
        assert abs(g.neg_mse_score(Xd, y) - blob["neg_mse"][i]) < 1e-6, kw
        if kind in ("lasso", "enet") and blob["n_iter"][i] > 0:
            n_iter_mismatch += int(g.model.n_iter_ != blob["n_iter"][i])
    assert n_iter_mismatch == 0


def test_predict_score_residuals_vs_numpy():
    rng = np.random.default_rng(21)
    X = rng.standard_normal((1501, 33))
    y = rng.standard_normal(1501)
    g = sglm.GLM("Gaussian", alpha=0.01, l1_ratio=0.5)
    g.fit(X, y)
    pred = X @ g.coef_ + g.intercept_
    assert np.allclose(g.predict(X), pred, rtol=0, atol=1e-12)
    assert np.allclose(g.predict(pd.DataFrame(X)), pred, rtol=0, atol=1e-12)
    r, mr = g.get_residuals(X, y)
    assert np.allclose(r, y - pred, atol=1e-12) and np.allclose(mr, y - y.mean(), atol=1e-12)
    assert abs(g.neg_mse_score(X, y) + np.mean((y - pred) ** 2)) < 1e-12
    assert abs(g.r2_score(X, y) - orc.r2_score(y, pred)) < 1e-12
    assert abs(sglm.calc_R2(r, mr) - orc.calc_R2(y - pred, y - y.mean())) < 1e-12


def test_warm_start_matches_oracle():
    rng = np.random.default_rng(31)
    X = rng.standard_normal((2000, 25))
    y = X[:, :5].sum(1) + rng.standard_normal(2000)
    beta = rng.standard_normal(25) * 0.1
    w_o, b_o, info = orc.enet_fit(X, y, 0.05, 0.5, True, 1000, 1e-4, coef_init=beta, use_gram=True)
    g = sglm.GLM("Gaussian", beta_=beta, beta0_=0.0, alpha=0.05, l1_ratio=0.5)
    g.fit(X, y)
    assert g.model.n_iter_ == info["n_iter"]
    assert coef_rel_err(g.coef_, w_o) < 1e-9


def test_error_behaviour():
    X = np.random.default_rng(0).standard_normal((50, 3))
    y = np.arange(50.0)
    Xn = X.copy()
    Xn[3, 1] = np.nan
    with pytest.raises(ValueError):
        sglm.GLM("Gaussian", alpha=0.1, l1_ratio=0.5).fit(Xn, y)
    with pytest.raises(ValueError):
        sglm.GLM("Gaussian", alpha=0.1, l1_ratio=0.5).fit(X, y[:-1])
    with pytest.raises(KeyError):
        sglm.GLM("Gaussian", alpha=0)
    with pytest.raises(TypeError):
        sglm.GLM("Gaussian", reg_lambda=0.1)                    # backend/test/test_sglm.py:25 (stale kwarg)
    with pytest.raises(NameError):
        sglm.GLM("NoSuchFamily")


# ------------------------------------------------------------------ Poisson
def test_poisson_vs_sklearn_optimum_golden():
    blob, meta = load_golden("poisson_ref")
    Xd, y = _design(blob), blob["y"]
    for i, kw in enumerate(meta["grid"]):
        g = sglm.GLM("Poisson", **dict(kw))
        g.fit(Xd, y)                                    # the reference raises AttributeError here
        assert coef_rel_err(g.coef_, blob["coefs"][i]) < 1e-6, (kw, coef_rel_err(g.coef_, blob["coefs"][i]))
        assert abs(g.intercept_ - blob["intercepts"][i]) < 1e-7, kw
        assert abs(g.r2_score(Xd, y) - blob["d2"][i]) < 1e-6, kw
        mu = np.exp(Xd @ g.coef_ + g.intercept_)
        assert np.allclose(g.predict(Xd), mu, rtol=1e-12)
        assert abs(g.neg_mse_score(Xd, y) + np.mean((y - mu) ** 2)) < 1e-10


# ------------------------------------------------------------------ CV grid
def test_cv_grid_vs_reference_golden():
    blob, meta = load_golden("cv_ref")
    Xd, y = _design(blob), blob["y"]
    cv_idx = [(blob[f"train{k}"], blob[f"test{k}"]) for k in range(meta["n_folds"])]
    for run in meta["runs"]:
        tag = run["tag"]
        kw_lst = [dict(k) for k in run["kwargs"]]
        res = sglm_cv.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", kw_lst, score_method=run["score_method"])
        assert res["best_params"] == run["best_params"], tag
        assert abs(res["best_score"] - run["best_score"]) < 1e-6, tag
        assert abs(res["best_score_std"] - run["best_score_std"]) < 1e-6, tag
        assert all("roll" not in k and "model_name" not in k for k in kw_lst)     # popped in place
        for j, r in enumerate(res["full_cv_results"]):
            assert r["glm_kwargs"] == run["result_kwargs"][j]
            assert list(r)[:11] == ['cv_coefs', 'cv_intercepts', 'cv_scores_train', 'cv_scores_test',
                                    'cv_mean_score_train', 'cv_mean_score', 'cv_std_score', 'cv_R2_score',
                                    'cv_mse_score', 'glm_kwargs', 'model']
            for k in range(meta["n_folds"]):
                assert coef_rel_err(r["cv_coefs"][:, k], blob[f"{tag}_coefs{j}"][:, k]) < 1e-4, (tag, j, k)
            assert np.allclose(r["cv_intercepts"], blob[f"{tag}_icpt{j}"], atol=1e-6)
            assert np.allclose(r["cv_scores_train"], blob[f"{tag}_tr{j}"], atol=1e-6), (tag, j)
            assert np.allclose(r["cv_scores_test"], blob[f"{tag}_te{j}"], atol=1e-6), (tag, j)
            agg = [r["cv_mean_score_train"], r["cv_mean_score"], r["cv_std_score"], r["cv_R2_score"],
                   r["cv_mse_score"]]
            assert np.allclose(agg, blob[f"{tag}_agg{j}"], atol=1e-6), (tag, j)
            assert coef_rel_err(r["model"].coef_, blob[f"{tag}_fullcoef{j}"]) < 1e-4
            assert abs(r["model"].intercept_ - float(blob[f"{tag}_fullicpt{j}"])) < 1e-6
            # the returned model object predicts and scores
            assert r["model"].predict(Xd[:7]).shape == (7,)


def test_cv_grid_vs_oracle_overlapping_and_repeated_indices():
    """Train sets that are not the complement of the test sets and repeated rows."""
    rng = np.random.default_rng(77)
    X0 = orc.synth_base(1500, 4, 77)
    shifts = [0, -2, -1, 1, 2]
    Xd = orc.timeshift_multiple(X0, shift_amt_list=shifts)[2:-2]
    y = orc.synth_response(Xd, orc.synth_kernels(4, shifts, 77), 77)
    n = Xd.shape[0]
    cv_idx = []
    for _ in range(3):
        perm = rng.permutation(n)
        train = np.concatenate([perm[:900], perm[:100]])           # 100 rows twice
        test = perm[850:1300]                                      # overlaps train
        cv_idx.append((train, test))
    grid = [dict(alpha=0.01, l1_ratio=0.5, max_iter=1000, fit_intercept=True),
            dict(alpha=1.0, l1_ratio=0, max_iter=1000, fit_intercept=True),
            dict(alpha=0.1, l1_ratio=1, max_iter=1000, fit_intercept=False)]
    want = orc.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
    got = sglm_cv.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
    assert got["best_params"] == want["best_params"]
    for a, b in zip(got["full_cv_results"], want["full_cv_results"]):
        for k in range(3):
            assert coef_rel_err(a["cv_coefs"][:, k], b["cv_coefs"][:, k]) < 1e-4
        assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], atol=1e-6)
        assert np.allclose(a["cv_scores_train"], b["cv_scores_train"], atol=1e-6)
        assert abs(a["cv_R2_score"] - b["cv_R2_score"]) < 1e-6
        assert abs(a["cv_mse_score"] - b["cv_mse_score"]) < 1e-6


def test_cv_single_params_and_poisson_cv_vs_oracle():
    X0 = orc.synth_base(1200, 3, 8)
    shifts = [0, -1, 1]
    Xd = orc.timeshift_multiple(X0, shift_amt_list=shifts)[1:-1]
    beta = orc.synth_kernels(3, shifts, 8)
    y = orc.synth_response(Xd, beta, 8)
    cv_idx = orc.synth_folds(Xd.shape[0], 3, seed=8, group=100)
    kw = dict(alpha=0.01, l1_ratio=0.3, max_iter=500, roll=3)
    resp = []
    got = sglm_cv.cv_glm_single_params(Xd, y, cv_idx, "Gaussian", kw, resp_list=resp, score_method="mse")
    assert resp and resp[0] is got and "roll" not in kw
    want = orc.cv_glm_single_params(Xd, y, cv_idx, "Gaussian", dict(alpha=0.01, l1_ratio=0.3, max_iter=500, roll=3))
    assert np.allclose(got["cv_scores_test"], want["cv_scores_test"], atol=1e-6)
    assert coef_rel_err(got["model"].coef_, want["model"].coef_) < 1e-4
    yp = orc.synth_response(Xd, beta, 8, poisson=True)
    gp = sglm_cv.cv_glm_single_params(Xd, yp, cv_idx, "Poisson", dict(alpha=0.01), score_method="r2", resp_list=[])
    wp = orc.cv_glm_single_params(Xd, yp, cv_idx, "Poisson", dict(alpha=0.01), score_method="r2")
    for k in range(3):
        assert coef_rel_err(gp["cv_coefs"][:, k], wp["cv_coefs"][:, k]) < 1e-6
    assert np.allclose(gp["cv_scores_test"], wp["cv_scores_test"], atol=1e-6)
    assert abs(gp["cv_R2_score"] - wp["cv_R2_score"]) < 1e-6


def test_reference_integration_flow_repaired():
    """backend/test/test_sglm_ez.py:19-74 with its stale unpacking repaired (5-tuple)."""
    X_tmp = pd.DataFrame(np.arange(200).reshape((100, 2)), columns=['A', 'B'])
    X_tmp['B'] = (X_tmp['B'] - 1) * 2 + 1
    X_tmp = sglm_ez.timeshift_cols(X_tmp, ['A'], pos_order=2)
    assert list(X_tmp.columns) == ['A', 'B', 'A_1', 'A_2']
    X_tmp = sglm_ez.diff_cols(X_tmp, ['A', 'B']).dropna()
    glm = sglm_ez.fit_GLM(X_tmp[['A']], X_tmp['B'], alpha=0.1)
    ref = orc.GLM("Gaussian", alpha=0.1).fit(X_tmp[['A']].values, X_tmp['B'].values)
    assert coef_rel_err(glm.coef_, ref.coef_) < 1e-4
    np.random.seed(0)
    cv_idx = sglm_ez.cv_idx_by_timeframe(X_tmp, y=None, num_folds=None, timesteps_per_bucket=20)
    lst = sglm_cv.generate_mult_params({'alpha': reversed([0.1, 1.0, 10.0]), 'l1_ratio': [0.1, 0.5, 0.9],
                                        'fit_intercept': [True, False]}, {'max_iter': 10000})
    assert len(lst) == 18
    best_score, best_std, best_params, best_model, cv_results = sglm_ez.simple_cv_fit(
        X_tmp[['A']], X_tmp['B'], cv_idx, lst, model_type='Normal')
    want = orc.cv_glm_mult_params(X_tmp[['A']].values, X_tmp['B'].values, cv_idx, "Gaussian",
                                  sglm_cv.generate_mult_params({'alpha': reversed([0.1, 1.0, 10.0]),
                                                                'l1_ratio': [0.1, 0.5, 0.9],
                                                                'fit_intercept': [True, False]}, {'max_iter': 10000}))
    assert best_params == want["best_params"]
    assert abs(best_score - want["best_score"]) < 1e-6 * max(1.0, abs(want["best_score"]))
    ols = orc.GLM("Gaussian", alpha=0, l1_ratio=0, max_iter=1).fit(X_tmp[['A']].values, X_tmp['B'].values)
    assert abs(ols.intercept_ - best_model.intercept_) < 0.01
    assert np.all(np.abs(ols.coef_ - best_model.coef_) < 0.01)
    glm2, hs, hm = sglm_ez.training_fit_holdout_score(X_tmp[['A']], X_tmp['B'], X_tmp[['A']], X_tmp['B'], best_params)
    assert hs > 0.99 and hm <= 0.0


def test_train_stats_identity_at_scale():
    """Size-independent property at a larger size: G(full) == G(test) + G(train) and the CV
    scores computed from statistics match an explicit pass over X."""
    T, C = 60_000, 96
    X = torch.randn((T, C), dtype=torch.float64, device="cuda")
    y = torch.randn((T,), dtype=torch.float64, device="cuda")
    m = (torch.rand(T, device="cuda") < 0.2).double()
    W = torch.stack([torch.ones_like(m), m, 1.0 - m])
    G = eng.suffstats(X, y[:, None], W, [T, float(m.sum()), float(T - m.sum())])
    err = (G[0] - G[1] - G[2]).abs().max().item()
    assert err <= 1e-9 * G[0].abs().max().item()
    Xh, yh = X.cpu().numpy(), y.cpu().numpy()
    test = np.flatnonzero(m.cpu().numpy() > 0)
    train = np.flatnonzero(m.cpu().numpy() == 0)
    r = sglm_cv.cv_glm_single_params(X, y, [(train, test)], "Gaussian", dict(alpha=0.01, l1_ratio=0.5),
                                     score_method="r2", resp_list=[])
    w, b = r["cv_coefs"][:, 0], r["cv_intercepts"][0]
    assert abs(r["cv_scores_test"][0] - orc.r2_score(yh[test], Xh[test] @ w + b)) < 1e-9
    assert abs(r["cv_scores_train"][0] - orc.r2_score(yh[train], Xh[train] @ w + b)) < 1e-9


# ------------------------------------------------------------------ BASELINE.json configurations
def _session(T, P, h_lo, h_hi, seed, poisson=False):
    import synth_data
    shifts = [0] + [s for s in range(-h_lo, h_hi + 1) if s != 0]
    X0 = synth_data.synth_base(T, P, seed)
    Xd = orc.timeshift_multiple(X0, shift_amt_list=shifts)
    Xd = Xd[~np.isnan(Xd).any(axis=1)]
    y = synth_data.synth_response(Xd, synth_data.synth_kernels(P, shifts, seed), seed, poisson=poisson)
    return X0, shifts, Xd, y


def test_config1_single_elasticnet_fit_vs_sklearn():
    """configs[0] shape (10 predictors x 41 shifts) at reduced T, against scikit-learn itself."""
    from sklearn.linear_model import ElasticNet
    X0, shifts, Xd, y = _session(20_000, 10, 20, 20, 101)
    assert Xd.shape[1] == 410
    design = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)
    assert design[20:-20].tobytes() == Xd.tobytes()
    for alpha, l1 in [(0.01, 0.5), (0.001, 0.9)]:
        ref = ElasticNet(alpha=alpha, l1_ratio=l1).fit(Xd, y)
        g = sglm.GLM("Gaussian", alpha=alpha, l1_ratio=l1)
        g.fit(Xd, y)
        assert g.model.n_iter_ == ref.n_iter_
        assert coef_rel_err(g.coef_, ref.coef_) < 1e-4
        assert abs(g.intercept_ - ref.intercept_) < 1e-8
        assert abs(g.r2_score(Xd, y) - ref.score(Xd, y)) < 1e-6


def test_config2_ridge_cv_grid_vs_oracle():
    """configs[1] shape (20 predictors x 61 shifts = 1220 columns, 5 folds) at reduced T."""
    import synth_data
    X0, shifts, Xd, y = _session(12_000, 20, 30, 30, 202)
    assert Xd.shape[1] == 1220
    cv_idx = synth_data.synth_folds(Xd.shape[0], 5, 202, group=500)
    grid = [dict(alpha=float(a), l1_ratio=0, max_iter=1000, fit_intercept=True) for a in np.logspace(-3, 3, 4)]
    want = orc.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
    for mode in ("dmma", "tc"):
        import os
        os.environ["SGLM_GRAM"] = mode
        try:
            got = sglm_cv.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
        finally:
            os.environ.pop("SGLM_GRAM", None)
        assert got["best_params"] == want["best_params"], mode
        for a, b in zip(got["full_cv_results"], want["full_cv_results"]):
            for k in range(5):
                assert coef_rel_err(a["cv_coefs"][:, k], b["cv_coefs"][:, k]) < 1e-7, mode
            assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], atol=1e-6), mode
            assert np.allclose(a["cv_scores_train"], b["cv_scores_train"], atol=1e-6), mode
            assert abs(a["cv_R2_score"] - b["cv_R2_score"]) < 1e-6
            assert coef_rel_err(a["model"].coef_, b["model"].coef_) < 1e-7


def test_config3_elasticnet_cv_grid_vs_oracle():
    """configs[2] shape (50 shifts, > 1024 columns, random overlapping folds) at reduced T: the whole wide-design
    path — Gram over the disjoint cells of the row sets, coordinate descent as cluster + per-model launch plan,
    row-split quadratic forms — against the oracle's CV grid (scikit-learn fits)."""
    import synth_data
    X0, shifts, Xd, y = _session(5_000, 24, 20, 29, 303)
    assert Xd.shape[1] == 1200 and len(shifts) == 50
    cv_idx = synth_data.synth_folds(Xd.shape[0], 3, 303, group=250)
    grid = [dict(alpha=float(a), l1_ratio=float(l), max_iter=1000, fit_intercept=True, tol=1e-4)
            for l in (0.2, 0.9) for a in np.logspace(-2.5, -0.5, 3)]
    want = orc.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2",
                                  engine="sklearn")
    got = sglm_cv.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
    assert nat.last_tc_plan["cells"]                                       # overlapping folds went through the cells
    assert eng._cd_plan(1200, len(grid) * 4, 4)[0][3] >= 2                 # ... and the heavy models through clusters
    assert got["best_params"] == want["best_params"]
    assert abs(got["best_score"] - want["best_score"]) < 1e-6
    for a, b in zip(got["full_cv_results"], want["full_cv_results"]):
        for k in range(3):
            assert coef_rel_err(a["cv_coefs"][:, k], b["cv_coefs"][:, k]) < 1e-4
        assert np.allclose(a["cv_intercepts"], b["cv_intercepts"], atol=1e-8)
        assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], atol=1e-6)
        assert np.allclose(a["cv_scores_train"], b["cv_scores_train"], atol=1e-6)
        assert abs(a["cv_R2_score"] - b["cv_R2_score"]) < 1e-6
        assert abs(a["cv_mse_score"] - b["cv_mse_score"]) < 1e-6
        assert coef_rel_err(a["model"].coef_, b["model"].coef_) < 1e-4
        assert a["model"].model.n_iter_ == b["model"].model.n_iter_


def test_config4_poisson_alpha_sweep_vs_sklearn_optimum():
    """configs[3] shape (20 predictors x 40 shifts) at reduced T, against TweedieRegressor driven
    to its optimum (newton-cholesky, tol 1e-12)."""
    from sklearn.linear_model import TweedieRegressor
    X0, shifts, Xd, y = _session(15_000, 20, 20, 19, 404, poisson=True)
    assert Xd.shape[1] == 800
    for alpha in (1e-2, 1.0):
        ref = TweedieRegressor(power=1, alpha=alpha, solver="newton-cholesky", tol=1e-12, max_iter=1000).fit(Xd, y)
        g = sglm.GLM("Poisson", alpha=alpha)
        g.fit(Xd, y)
        assert coef_rel_err(g.coef_, ref.coef_) < 1e-6
        assert abs(g.intercept_ - ref.intercept_) < 1e-7
        assert abs(g.r2_score(Xd, y) - ref.score(Xd, y)) < 1e-6


def test_poisson_newton_with_tensor_core_hessian_reaches_the_same_optimum():
    """The large-problem Poisson path: Newton steps whose Hessian X'WX is the tcgen05 digit-plane Gram of
    sqrt(W) X truncated to 4 planes (approximate) and whose gradient X'(mu - y) is exact.  Same optimum as
    TweedieRegressor driven to convergence and as the exact fp64 IRLS path, with and without fold weights."""
    from sklearn.linear_model import TweedieRegressor
    X0, shifts, Xd, y = _session(15_000, 20, 20, 19, 404, poisson=True)
    Xg, yg = torch.from_numpy(Xd).cuda(), torch.from_numpy(y).cuda()
    rng = np.random.default_rng(1)
    rw_h = (rng.random(Xd.shape[0]) < 0.8).astype(np.float64)
    rw = torch.from_numpy(rw_h).cuda()
    old = eng.POISSON_TC
    try:
        for alpha, weights in [(1e-2, None), (1.0, None), (1e-2, rw)]:
            eng.POISSON_TC = True
            w_tc, b_tc, n_tc = eng.poisson_irls(Xg, yg, alpha, True, weights)
            eng.POISSON_TC = False
            w_ex, b_ex, n_ex = eng.poisson_irls(Xg, yg, alpha, True, weights)
            assert coef_rel_err(w_tc, w_ex) < 1e-7 and abs(b_tc - b_ex) < 1e-8
            assert n_tc <= n_ex + 3
            sel = slice(None) if weights is None else rw_h > 0
            ref = TweedieRegressor(power=1, alpha=alpha, solver="newton-cholesky", tol=1e-12, max_iter=1000).fit(Xd[sel], y[sel])
            assert coef_rel_err(w_tc, ref.coef_) < 1e-6
            assert abs(b_tc - ref.intercept_) < 1e-7
        eng.POISSON_TC = True
        w0, b0, _ = eng.poisson_irls(Xg, yg, 0.1, False, None)               # no intercept
        ref = TweedieRegressor(power=1, alpha=0.1, fit_intercept=False, solver="newton-cholesky", tol=1e-12,
                               max_iter=1000).fit(Xd, y)
        assert coef_rel_err(w0, ref.coef_) < 1e-6 and b0 == 0.0
    finally:
        eng.POISSON_TC = old


def test_full_size_properties_config2_and_3():
    """Size-independent properties at (near) BASELINE sizes, checked with an independent fp64
    implementation (torch): Ridge normal equations hold; ElasticNet models satisfy the duality-gap
    stopping rule recomputed from explicit residuals; selection equals the argmax of the scores."""
    import synth_data
    T, P = 400_000, 20
    shifts = [0] + [s for s in range(-30, 31) if s != 0]
    X0 = torch.from_numpy(synth_data.synth_base(T, P, 303)).cuda()
    X = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[30:T - 30]          # 399,940 x 1220 (3.9 GB)
    n, C = X.shape
    beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 303)).cuda()
    y = X @ beta + 2.0 * torch.randn(n, dtype=torch.float64, device="cuda")
    y = (y - y.mean()) / y.std()
    cv_idx = synth_data.synth_folds(n, 3, 303, group=1000)
    grid = [dict(alpha=1.0, l1_ratio=0, max_iter=100), dict(alpha=100.0, l1_ratio=0, max_iter=100),
            dict(alpha=1e-3, l1_ratio=0.5, max_iter=1000), dict(alpha=1e-2, l1_ratio=0.9, max_iter=1000)]
    res = sglm_cv.cv_glm_mult_params(X, y, cv_idx, "Gaussian", [dict(g) for g in grid], score_method="r2")
    scores = [r["cv_R2_score"] for r in res["full_cv_results"]]
    assert res["best_score"] == max(scores) and res["best_params"] == res["full_cv_results"][int(np.argmax(scores))]["glm_kwargs"]
    Xc = X - X.mean(0)
    yc = y - y.mean()
    for r in res["full_cv_results"]:
        kw, w = r["glm_kwargs"], torch.from_numpy(r["model"].coef_).cuda()
        resid = yc - Xc @ w
        if kw["l1_ratio"] == 0:                                    # (Xc'Xc + alpha I) w = Xc'yc
            g = Xc.T @ resid - kw["alpha"] * w
            assert g.abs().max().item() <= 1e-8 * (Xc.T @ yc).abs().max().item()
        else:                                                      # sklearn's stopping rule: gap <= tol * ||yc||^2
            l1 = kw["alpha"] * kw["l1_ratio"] * n
            l2 = kw["alpha"] * (1 - kw["l1_ratio"]) * n
            xta = Xc.T @ resid - l2 * w
            dn = xta.abs().max().item()
            R2, ww = float(resid @ resid), float(w @ w)
            scale = min(1.0, l1 / dn)
            gap = 0.5 * (R2 + l2 * ww) + l1 * float(w.abs().sum()) - (-0.5 * scale ** 2 * (R2 + l2 * ww) + scale * float(resid @ yc))
            assert gap <= 1e-4 * float(yc @ yc) * (1 + 1e-6)
        b = float(y.mean() - X.mean(0) @ w)
        assert abs(b - r["model"].intercept_) < 1e-9
    for f, (tr, te) in enumerate(cv_idx):                          # fold scores from statistics == explicit pass
        r = res["full_cv_results"][2]
        w = torch.from_numpy(r["cv_coefs"][:, f]).cuda()
        te_t = torch.from_numpy(te).cuda()
        pred = X[te_t] @ w + r["cv_intercepts"][f]
        yt = y[te_t]
        r2 = 1.0 - float(((yt - pred) ** 2).sum() / ((yt - yt.mean()) ** 2).sum())
        assert abs(r2 - r["cv_scores_test"][f]) < 1e-8


def test_second_generation_lag_builder_vs_reference_golden():
    """`timeshift_vals_by_dict` (sglm/sglm/features/setup_model_fit.py:43-96): predictor-major layout."""
    import setup_model_fit
    blob, meta = load_golden("by_dict_ref")
    df = pd.DataFrame(blob["df"], columns=meta["columns"])
    for i, case in enumerate(meta["cases"]):
        d = {k: tuple(v) for k, v in case["d"].items()}
        out, names = setup_model_fit.timeshift_vals_by_dict(df, d, keep_nans=case["keep_nans"])
        assert names == case["names"] and list(out.columns) == case["out_columns"]
        assert np.array_equal(np.asarray(out.index), blob[f"idx{i}"])
        assert out.to_numpy(dtype=np.float64).tobytes() == blob[f"out{i}"].tobytes()
    got = setup_model_fit.X_cols_dict_to_default({"a": (0, 0), "b": None, "c": (-1, 2)})
    assert {k: list(v) for k, v in got.items()} == meta["default"]


def test_ols_rank_deficient_matches_lstsq_min_norm():
    """alpha == 0 on a rank-deficient design (an all-zero column and a duplicated column):
    scikit-learn's LinearRegression returns the minimum-norm least-squares solution."""
    from sklearn.linear_model import LinearRegression
    rng = np.random.default_rng(12)
    X = rng.standard_normal((500, 12))
    X[:, 3] = 0.0
    X[:, 7] = X[:, 2]
    y = X @ rng.standard_normal(12) + 0.1 * rng.standard_normal(500)
    ref = LinearRegression().fit(X, y)
    g = sglm.GLM("Gaussian", alpha=0, l1_ratio=0.5, max_iter=10)
    g.fit(X, y)
    assert coef_rel_err(g.coef_, ref.coef_) < 1e-7
    assert abs(g.intercept_ - ref.intercept_) < 1e-9
    assert np.allclose(g.predict(X), ref.predict(X), atol=1e-9)


def test_warm_started_paths_reach_the_same_optimum():
    """Opt-in warm-started alpha paths: same optimum as the cold-start (reference) mode —
    <= 1e-8 on coefficients at tol = 1e-10, and within solver tolerance at 1e-4."""
    import synth_data
    X0, shifts, Xd, y = _session(6000, 6, 4, 4, 55)
    cv_idx = synth_data.synth_folds(Xd.shape[0], 3, 55, group=200)
    def run(tol, warm):
        grid = [dict(alpha=float(a), l1_ratio=l, max_iter=5000, tol=tol) for l in (0.2, 0.8) for a in np.logspace(-3, -1, 6)]
        eng.WARM_START_PATHS = warm
        try:
            return sglm_cv.cv_glm_mult_params(Xd, y, cv_idx, "Gaussian", grid, score_method="r2")
        finally:
            eng.WARM_START_PATHS = False
    cold, warm = run(1e-10, False), run(1e-10, True)
    assert cold["best_params"] == warm["best_params"]
    sweeps = lambda res: sum(r["_fit_info"]["cd_info"][:, 2].sum() for r in res["full_cv_results"])
    for a, b in zip(cold["full_cv_results"], warm["full_cv_results"]):
        assert a["glm_kwargs"] == b["glm_kwargs"]
        for k in range(3):
            assert coef_rel_err(b["cv_coefs"][:, k], a["cv_coefs"][:, k]) < 1e-8
        assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], atol=1e-9)
    assert sweeps(warm) < sweeps(cold)
    cold4, warm4 = run(1e-4, False), run(1e-4, True)
    for a, b in zip(cold4["full_cv_results"], warm4["full_cv_results"]):
        assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], atol=1e-3)


def test_second_generation_closed_form_fit():
    """`glm.closed_form = True` (sglm/sglm/models/sglm.py:263-293): lstsq on [X | 1]."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((400, 9))
    y = X @ rng.standard_normal(9) + 2.5 + 0.1 * rng.standard_normal(400)
    want = np.linalg.lstsq(np.concatenate([X, np.ones((400, 1))], axis=1), y, rcond=-1)[0]
    g = sglm.GLM("Gaussian", alpha=0.1, l1_ratio=0.5, fit_intercept=True)
    g.closed_form = True
    g.fit(X, y)
    assert coef_rel_err(np.append(g.coef_, g.intercept_), want) < 1e-9
    assert np.allclose(g.predict(X), np.concatenate([X, np.ones((400, 1))], axis=1) @ want, atol=1e-9)
    g2 = sglm.fit_GLM(pd.DataFrame(X), pd.Series(y), alpha=0.01)
    assert g2.coef_.shape == (9,)


def test_fit_set_in_place_semantics_from_threads():
    """GLM.fit_set (backend/sglm.py:254-312) writes into caller-owned arrays and lists; the
    reference calls it from 4 Python threads (backend/sglm_cv.py:162-170)."""
    import threading
    rng = np.random.default_rng(8)
    X = rng.standard_normal((900, 14))
    y = X[:, :3].sum(1) + rng.standard_normal(900)
    folds = [(np.r_[0:600], np.r_[600:900]), (np.r_[300:900], np.r_[0:300]), (np.r_[0:300, 600:900], np.r_[300:600]),
             (np.r_[100:800], np.r_[0:100, 800:900])]
    n = len(folds)
    cv_coefs, cv_icpt = np.zeros((14, n)), np.zeros(n)
    tr, te = np.zeros(n), np.zeros(n)
    resids, mean_resids = [], []
    def work(k):
        a, b = folds[k]
        g = sglm.GLM("Gaussian", alpha=0.02, l1_ratio=0.5, score_method="r2")
        g.fit_set(X[a], y[a], X[b], y[b], cv_coefs, cv_icpt, tr, te, k, resids=resids, mean_resids=mean_resids, id_fit=k)
    threads = [threading.Thread(target=work, args=(k,)) for k in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(resids) == n and len(mean_resids) == n
    for k, (a, b) in enumerate(folds):
        ref = orc.GLM("Gaussian", alpha=0.02, l1_ratio=0.5, score_method="r2").fit(X[a], y[a])
        assert coef_rel_err(cv_coefs[:, k], ref.coef_) < 1e-4
        assert abs(cv_icpt[k] - ref.intercept_) < 1e-8
        assert abs(tr[k] - ref.r2_score(X[a], y[a])) < 1e-6 and abs(te[k] - ref.r2_score(X[b], y[b])) < 1e-6
    pooled = sglm.calc_R2(np.concatenate(resids), np.concatenate(mean_resids))
    assert -1.0 < pooled < 1.0


def test_quadform_gemm_path_equals_row_streaming_kernel():
    """v' A v of many vectors through the fp64 GEMM (sglm_quadform_gemm_f64) against the row-streaming kernel and numpy:
    ragged sizes (n not a multiple of the tiles, M not a multiple of 64), padded leading dimensions."""
    rng = np.random.default_rng(11)
    for n, M in [(203, 130), (1221, 300), (64, 128)]:
        B = rng.standard_normal((n, n))
        A = torch.from_numpy(B @ B.T).cuda()
        Apad = torch.zeros((n, n + 5), dtype=torch.float64, device="cuda")
        Apad[:, :n] = A
        V = torch.from_numpy(rng.standard_normal((M, n + (n & 1)))).cuda()
        want = np.einsum("mi,ij,mj->m", V[:, :n].cpu().numpy(), A.cpu().numpy(), V[:, :n].cpu().numpy())
        old = eng.QUADFORM_GEMM_MIN
        try:
            eng.QUADFORM_GEMM_MIN = 1 << 30
            q_rows = eng.quadform(Apad[:, :n], V).cpu().numpy()
            eng.QUADFORM_GEMM_MIN = 1
            q_gemm = eng.quadform(Apad[:, :n], V).cpu().numpy()
        finally:
            eng.QUADFORM_GEMM_MIN = old
        scale = np.abs(want).max()
        assert np.max(np.abs(q_rows - want)) < 1e-12 * scale
        assert np.max(np.abs(q_gemm - want)) < 1e-12 * scale
