"""GPU tests of the cluster coordinate-descent kernel (csrc/cd_cluster.cuh): for every supported
(models per cluster x CTAs per cluster) shape and for multi-part launch plans the coefficients,
sweep counts and update counts must be IDENTICAL (bit for bit) to the first-generation
one-CTA-per-model kernel, which in turn is pinned iterate-for-iterate against the oracle
(test_gpu_parity.py::test_cd_kernel_matches_oracle_gram_cd_iterate_for_iterate)."""
import numpy as np
import pytest

from conftest import coef_rel_err
from oracle import sglm_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import _engine as eng  # noqa: E402


def _problems(n, C, n_sets, seed, zero_col=None):
    """n_sets centred problems (full data + folds as complements) from a correlated random design."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, C))
    X[:, 1:] += 0.6 * X[:, :-1]                       # neighbouring columns correlated (lag-like)
    if zero_col is not None:
        X[:, zero_col] = 0.0                          # Q[j,j] == 0 branch
    beta = rng.standard_normal(C) * (rng.random(C) < 0.3)
    y = X @ beta + 2.0 * rng.standard_normal(n)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    W = [torch.ones(n, dtype=torch.float64, device="cuda")]
    sizes = [n]
    for f in range(n_sets - 1):
        idx = np.arange(f, n, n_sets + 1)
        W.append(eng.index_counts(idx, n))
        sizes.append(len(idx))
    G = eng.suffstats(Xd, yd[:, None].contiguous(), torch.stack(W), sizes)
    probs = [eng.center(G[0], None, C, 1, 0, True)]
    probs += [eng.center(G[0], G[1 + f], C, 1, 0, True) for f in range(n_sets - 1)]
    eng.fetch_scalars(probs)
    return probs, X, y


def _solve(models, C, plan, **kw):
    old = eng.CD_PLAN
    eng.CD_PLAN = plan
    try:
        W, info, status = eng.solve_models(models, C, **kw)
        torch.cuda.synchronize()
        return W.cpu().numpy(), info, status
    finally:
        eng.CD_PLAN = old


SHAPES = ["1x1", "1x2", "1x4", "1x8", "2x1", "2x2", "2x4", "2x8", "4x1", "4x2", "4x4", "4x8", "8x2", "8x4", "8x8"]


@pytest.mark.parametrize("C,n_sets", [(410, 3), (333, 2), (96, 2)])
def test_every_shape_is_bit_identical_to_the_first_generation_kernel(C, n_sets):
    probs, _, _ = _problems(3000, C, n_sets, seed=100 + C, zero_col=7)
    grid = [(0.5, 0.5, 1e-4, 1000), (0.05, 0.5, 1e-4, 1000), (0.005, 0.1, 1e-4, 1000), (0.02, 1.0, 1e-4, 1000),
            (1e-3, 0.9, 1e-8, 1000), (50.0, 0.5, 1e-4, 1000), (0.01, 0.3, 1e-4, 2)]
    models = [eng.ModelSpec(p, "enet", a, l, mi, tol) for (a, l, tol, mi) in grid for p in probs]   # 7 per problem: ragged groups
    W0, i0, s0 = _solve(models, C, "0x0")
    assert i0[:, 3].sum() > 0 and (i0[:, 2] == 0).any() and (i0[:, 2] == 2).any()      # incl. "converged at w = 0" and max_iter hits
    for shape in SHAPES:
        if shape.startswith("8x") and (C + 31) // 32 < 4:
            continue                                    # groups of 8 need at least a 2-CTA cluster
        W1, i1, s1 = _solve(models, C, shape)
        assert np.array_equal(W0, W1), shape
        assert np.array_equal(i0[:, 2], i1[:, 2]), (shape, "n_iter")
        assert np.array_equal(i0[:, 3], i1[:, 3]), (shape, "row updates")
        assert np.array_equal(s0, s1), shape
        assert np.allclose(i0[:, 0], i1[:, 0], rtol=1e-6, atol=1e-9 * np.abs(i0[:, 1]).max()), (shape, "gap")


def test_multi_part_plans_and_screening_off():
    C = 410
    probs, _, _ = _problems(2500, C, 3, seed=5)
    models = [eng.ModelSpec(p, "enet", a, l, 1000, 1e-4) for l in (0.2, 0.8) for a in np.logspace(-3, 0, 6) for p in probs]
    W0, i0, _ = _solve(models, C, "0x0")
    for plan in ["4x2@0.3,0x0", "2x4@0.25,4x1@0.25,1x1", "0x0@0.5,2x2"]:
        W1, i1, _ = _solve(models, C, plan)
        assert np.array_equal(W0, W1), plan
        assert np.array_equal(i0[:, 2:4], i1[:, 2:4]), plan
    Wn0, in0, _ = _solve(models, C, "0x0", do_screening=False)
    Wn1, in1, _ = _solve(models, C, "4x2", do_screening=False)
    assert np.array_equal(Wn0, Wn1) and np.array_equal(in0[:, 2:4], in1[:, 2:4])
    assert coef_rel_err(Wn0[:, :C], W0[:, :C]) < 1e-4            # screening does not change the optimum


def test_warm_start_through_the_cluster_kernel():
    C = 200
    probs, _, _ = _problems(2000, C, 2, seed=9)
    rng = np.random.default_rng(3)
    inits = [rng.standard_normal(C) * 0.05 * (rng.random(C) < 0.5) for _ in range(6)]
    models = [eng.ModelSpec(probs[k % 2], "enet", 0.02 * (k + 1), 0.5, 1000, 1e-4, coef_init=inits[k]) for k in range(6)]
    W0, i0, _ = _solve(models, C, "0x0")
    for shape in ["2x2", "4x1", "1x4"]:
        W1, i1, _ = _solve(models, C, shape)
        assert np.array_equal(W0, W1), shape
        assert np.array_equal(i0[:, 2:4], i1[:, 2:4]), shape


def test_cluster_kernel_against_the_oracle_iterate_for_iterate():
    """Same check as the first-generation kernel's: sweep counts equal to the oracle's Gram CD
    (plain-C restatement of sklearn _cd_fast.pyx:1095-1290), coefficients to 1e-9."""
    C = 130
    probs, X, y = _problems(3000, C, 1, seed=21, zero_col=3)
    n = X.shape[0]
    for alpha, l1r, tol, mi in [(0.05, 0.5, 1e-4, 1000), (0.3, 1.0, 1e-4, 1000), (2e-3, 0.1, 1e-8, 1000)]:
        w_o, b_o, info_o = orc.enet_fit(X, y, alpha, l1r, True, mi, tol, use_gram=True)
        W1, i1, _ = _solve([eng.ModelSpec(probs[0], "enet", alpha, l1r, mi, tol)], C, "1x4")
        assert int(i1[0, 2]) == info_o["n_iter"], (alpha, l1r)
        assert coef_rel_err(W1[0, :C], w_o) < 1e-9


def test_wide_design_default_plan_matches_first_generation():
    """C = 2000 (the width of BASELINE configs[2]): the default plan (groups of 4 models on 2-CTA clusters
    for the heavy part) against the first-generation kernel, bit for bit."""
    C = 2000
    probs, _, _ = _problems(6000, C, 2, seed=77)
    models = [eng.ModelSpec(p, "enet", a, l, 300, 1e-4) for l in (0.1, 0.9) for a in np.logspace(-2.5, 0, 10) for p in probs]
    assert eng._cd_plan(C, len(models))[0][2] > 0                 # the default really is the cluster kernel
    W0, i0, _ = _solve(models, C, "0x0")
    W1, i1, _ = _solve(models, C, None)
    assert np.array_equal(W0, W1)
    assert np.array_equal(i0[:, 2:4], i1[:, 2:4])


def test_random_shapes_and_widths_against_the_first_generation_kernel():
    """Seeded sweep over design widths that are not multiples of 32 (odd widths, one or two blocks per CTA,
    a ragged last block), ragged groups and random launch plans."""
    rng = np.random.default_rng(2024)
    for case in range(8):
        C = int(rng.integers(40, 700))
        n_sets = int(rng.integers(1, 4))
        probs, _, _ = _problems(int(rng.integers(C + 200, 2500)), C, n_sets, seed=500 + case,
                                zero_col=int(rng.integers(0, C)) if case % 2 else None)
        n_models = int(rng.integers(1, 14))
        models = [eng.ModelSpec(probs[int(rng.integers(0, n_sets))], "enet", float(10 ** rng.uniform(-3.5, 0.5)),
                                float(rng.choice([0.05, 0.5, 1.0])), int(rng.choice([3, 1000])), float(rng.choice([1e-4, 1e-7])))
                  for _ in range(n_models)]
        W0, i0, s0 = _solve(models, C, "0x0")
        shapes = list(rng.choice([s for s in SHAPES if not s.startswith("8x") or C >= 128], size=3, replace=False))
        plans = shapes + [f"{shapes[0]}@0.4,{shapes[1]}", f"{shapes[2]}@0.5,0x0"]
        for plan in plans:
            W1, i1, s1 = _solve(models, C, plan)
            assert np.array_equal(W0, W1), (case, C, plan)
            assert np.array_equal(i0[:, 2:4], i1[:, 2:4]), (case, C, plan)
            assert np.array_equal(s0, s1), (case, C, plan)
