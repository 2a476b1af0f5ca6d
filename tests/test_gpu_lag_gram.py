"""GPU tests of the lag-design statistics (csrc/gram_tc.cu tc_expand_kernel, _engine.LagRecipe): the tensor-core Gram of
a lag design computed from the BASE SIGNALS' digit planes — the T x (P*L) design is never built — must equal the
Gram of the built design: bit for bit whenever both analyses give the same exponents / plane counts (integer plane
Grams are exact), to 1e-13 when a column's extreme value sits in the few rows its lagged copies do not all see.
The reference arithmetic replaced is the same X'X / X'y of every sklearn fit (backend/sglm.py:241) on the design of
backend/sglm_pp.py:58-103."""
import numpy as np
import pandas as pd
import pytest

from conftest import coef_rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import synth_data  # noqa: E402
import _engine as eng  # noqa: E402
import _sglm_native as nat  # noqa: E402
import sglm_cv  # noqa: E402
import sglm_pp  # noqa: E402


def _design(T, P, shifts, seed, shift_inx=[]):
    X0 = synth_data.synth_base(T, P, seed)
    d = sglm_pp.timeshift_multiple(X0, shift_inx=shift_inx, shift_amt_list=shifts, device=True).dropna()
    rec = d.lag_recipe()
    assert rec is not None and rec.window() is not None
    return X0, d, rec


@pytest.mark.parametrize("T,P,shifts,group", [
    (60_000, 12, [0] + [s for s in range(-7, 9) if s != 0], 500),      # both signs, trial-block folds (contiguous runs)
    (33_333, 7, [0, 1, 2, 3, 5, 8], 1),                                  # positive shifts only, single-row folds (no runs)
    (20_011, 5, [-6, -3, 0], 37),                                        # negative only, ragged sizes
])
def test_lag_statistics_equal_the_statistics_of_the_built_design(T, P, shifts, group):
    X0, d, rec = _design(T, P, shifts, seed=T % 97)
    Xd = rec.tensor()
    n, C = Xd.shape
    rng = np.random.default_rng(3)
    Yd = torch.from_numpy(rng.standard_normal((n, 2))).cuda().contiguous()
    folds = synth_data.synth_folds(n, 4, 3, group=group)
    rows = [None] + [torch.from_numpy(b).cuda() for _, b in folds]
    old = eng.TC_CELLS
    try:
        for cells in (None, False):
            eng.TC_CELLS = cells
            G0, s0 = eng.suffstats_tc(Xd, Yd, rows)
            assert not nat.last_tc_plan.get("lag")
            G1, s1 = eng.suffstats_tc(rec, Yd, rows)
            assert nat.last_tc_plan["lag"]
            torch.cuda.synchronize()
            assert np.array_equal(s0, s1)
            assert torch.equal(G0, G1), (cells, float((G0 - G1).abs().max()))
    finally:
        eng.TC_CELLS = old


def test_lag_statistics_when_the_extreme_value_sits_in_an_edge_row():
    """The largest |value| of a signal in the very first base row: only the most-delayed copies of the signal contain
    it, so the per-column analysis of the built design gives the other copies a smaller exponent than the base-signal
    analysis does.  Different digit planes, same (exact) integer Grams up to the fp64 recombination: 1e-13."""
    T, P = 25_000, 6
    shifts = [0] + [s for s in range(-4, 6) if s != 0]
    X0 = synth_data.synth_base(T, P, 11)
    X0[0, P - 2] = 1000.0 + 1.0 / 3.0
    X0[T - 1, 0] = 3.0                                    # a 0/1 column that needs more planes only in its last row
    d = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, device=True).dropna()
    rec = d.lag_recipe()
    Xd = rec.tensor()
    n = Xd.shape[0]
    Yd = torch.from_numpy(np.random.default_rng(0).standard_normal((n, 1))).cuda().contiguous()
    rows = [None, torch.arange(0, n, 3, device="cuda")]
    G0, s0 = eng.suffstats_tc(Xd, Yd, rows)
    G1, s1 = eng.suffstats_tc(rec, Yd, rows)
    torch.cuda.synchronize()
    scale = G0.abs().amax(dim=(1, 2), keepdim=True)
    assert float(((G0 - G1).abs() / scale).max()) < 1e-13
    W = torch.stack([torch.ones(n, dtype=torch.float64, device="cuda"), eng.index_counts(rows[1], n)])
    G_ref = eng.suffstats(Xd, Yd, W, [n, int(rows[1].numel())])
    assert float(((G_ref - G1).abs() / scale).max()) < 1e-12


def test_cv_grid_on_a_lazy_design_equals_the_grid_on_the_built_design(monkeypatch):
    """cv_glm_mult_params(DeviceDesign, ...) — statistics from the base signals — against the same call on the built
    CUDA tensor and on the pandas frame a reference user would hold: identical selection, coefficients bit-equal
    between the two device paths.  Also a column subset (d[x_cols]) and a design whose columns read outside the base
    signals (fill 0, no dropna: the recipe has no window and the design is built)."""
    monkeypatch.setenv("SGLM_GRAM", "tc")
    T, P = 12_000, 6
    shifts = [0, -2, -1, 1, 2, 3]
    X0 = synth_data.synth_base(T, P, 21)
    df = pd.DataFrame(X0, columns=[f"s{i}" for i in range(P)])
    lazy = sglm_pp.timeshift_multiple(df, shift_amt_list=shifts, device=True).dropna()
    host = sglm_pp.timeshift_multiple(df, shift_amt_list=shifts, device=False).dropna()
    assert list(lazy.columns) == list(host.columns) and len(lazy) == len(host)
    beta = synth_data.synth_kernels(P, shifts, 21)
    y = synth_data.synth_response(host.values, beta, 21)
    folds = synth_data.synth_folds(len(host), 3, 5, group=400)
    grid = [dict(alpha=a, l1_ratio=l, max_iter=500, fit_intercept=True) for l in (0.2, 0.9) for a in (1e-3, 1e-2, 0.1)]
    grid += [dict(alpha=1.0, l1_ratio=0.0, fit_intercept=True), dict(alpha=0.0, l1_ratio=0.5, max_iter=100, fit_intercept=True)]
    r_lazy = sglm_cv.cv_glm_mult_params(lazy, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
    assert nat.last_tc_plan["lag"]
    r_dev = sglm_cv.cv_glm_mult_params(torch.from_numpy(host.values).cuda(), torch.from_numpy(y).cuda(), folds, "Gaussian",
                                       [dict(g) for g in grid], score_method="r2")
    assert not nat.last_tc_plan.get("lag")
    assert r_lazy["best_params"] == r_dev["best_params"]
    for a, b in zip(r_lazy["full_cv_results"], r_dev["full_cv_results"]):
        assert np.array_equal(a["cv_coefs"], b["cv_coefs"])
        assert np.array_equal(a["model"].coef_, b["model"].coef_)
        assert np.allclose(a["cv_scores_test"], b["cv_scores_test"], rtol=0, atol=1e-12)
    # a column subset of the lazy design
    cols = [c for c in lazy.columns if not c.startswith("s0")]
    r_sub = sglm_cv.cv_glm_mult_params(lazy[cols], y, folds, "Gaussian", [dict(g) for g in grid[:3]], score_method="r2")
    assert nat.last_tc_plan["lag"]
    r_sub_h = sglm_cv.cv_glm_mult_params(torch.from_numpy(host[cols].values).cuda(), y, folds, "Gaussian",
                                         [dict(g) for g in grid[:3]], score_method="r2")
    for a, b in zip(r_sub["full_cv_results"], r_sub_h["full_cv_results"]):
        assert np.array_equal(a["cv_coefs"], b["cv_coefs"])
    # columns that read outside the base signals (zero fill, every row kept): no window -> the design is built
    lazy0 = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, fill_value=0.0, device=True)
    assert lazy0.lag_recipe().window() is None
    y0 = synth_data.synth_response(np.asarray(lazy0), beta, 3)
    folds0 = synth_data.synth_folds(T, 3, 5, group=400)
    r0 = sglm_cv.cv_glm_mult_params(lazy0, y0, folds0, "Gaussian", [dict(g) for g in grid[:2]], score_method="r2")
    assert not nat.last_tc_plan.get("lag")
    r0_h = sglm_cv.cv_glm_mult_params(torch.from_numpy(np.asarray(lazy0)).cuda(), y0, folds0, "Gaussian",
                                      [dict(g) for g in grid[:2]], score_method="r2")
    for a, b in zip(r0["full_cv_results"], r0_h["full_cv_results"]):
        assert np.array_equal(a["cv_coefs"], b["cv_coefs"])


def test_lag_statistics_reject_non_finite_base_signals():
    T, P = 9_000, 4
    X0 = synth_data.synth_base(T, P, 2)
    X0[4000, 1] = np.inf                                   # not NaN: dropna keeps the row, the analysis must refuse it
    d = sglm_pp.timeshift_multiple(X0, shift_amt_list=[0, 1, 2], device=True).dropna()
    rec = d.lag_recipe()
    Yd = torch.zeros((rec.shape[0], 1), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        eng.suffstats_tc(rec, Yd, [None])


def test_lag_statistics_at_scale_multi_segment_cells():
    """900k x 600 lag design (12 signals x 50 shifts), 5 trial-block folds: cells above one int32-safe segment, several K
    parts — bit-equal to the statistics of the built design."""
    T, P = 900_000, 12
    shifts = [0] + [s for s in range(-20, 30) if s != 0]
    X0, d, rec = _design(T, P, shifts, seed=4)
    Xd = rec.tensor()
    n = Xd.shape[0]
    Yd = torch.from_numpy(np.random.default_rng(1).standard_normal((n, 1))).cuda().contiguous()
    folds = synth_data.synth_folds(n, 5, 9)
    rows = [None] + [torch.from_numpy(b).cuda() for _, b in folds]
    G0, s0 = eng.suffstats_tc(Xd, Yd, rows)
    G1, s1 = eng.suffstats_tc(rec, Yd, rows)
    torch.cuda.synchronize()
    assert nat.last_tc_plan["lag"] and nat.last_tc_plan["cells"]
    assert np.array_equal(s0, s1) and torch.equal(G0, G1)
