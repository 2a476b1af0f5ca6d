"""torchrun --nproc-per-node N scripts/strong_check.py : sglm_dist.cv_grid_strong (one grid over N GPUs: row-sharded
tcgen05 Gram with exact int64 all-reduce, models dealt by cost) must return EXACTLY what the one-GPU path returns
(same statistics bits -> same iterates): best_params, coefficients, scores compared with ==."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sabatinilab-glm_b200")):
    sys.path.insert(0, p)
import synth_data  # noqa: E402
import sglm_cv, sglm_dist, sglm_pp  # noqa: E402,E401

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
T, P = int(os.environ.get("CHECK_T", 400_000)), 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
grid = [dict(alpha=float(a), l1_ratio=float(l), max_iter=1000, fit_intercept=True, tol=1e-4)
        for l in (0.1, 0.5, 0.9) for a in np.logspace(-4, 0, 8)] + [dict(alpha=10.0, l1_ratio=0.0, max_iter=10),
                                                                   dict(alpha=0.0, l1_ratio=0.0, max_iter=10),
                                                                   dict(alpha=0.01, l1_ratio=0.5, max_iter=1000, roll=7)]
X0 = y = folds = None
if rank == 0:
    X0 = synth_data.synth_base(T, P, 31)
    d = sglm_pp.timeshift_multiple(torch.from_numpy(X0).cuda(), shift_amt_list=shifts)[29:T - 20]
    beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 31)).cuda()
    s = d @ beta
    y = (s + 1.5 * float(s.std()) * torch.randn(d.shape[0], dtype=torch.float64, device="cuda")).cpu().numpy()
    y = (y - y.mean()) / y.std()
    folds = synth_data.synth_folds(d.shape[0], 5, 31, group=1000)
    single = sglm_cv.cv_glm_mult_params(d, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
    del d, s
    # the one-GPU grid on the LAZY design (statistics from the base signals' digit planes) must equal the grid on the
    # built design; the strong-scaled grid below takes the lazy path on every rank
    lazy = sglm_pp.timeshift_multiple(torch.from_numpy(X0).cuda(), shift_amt_list=shifts, device=True).dropna()
    single_lazy = sglm_cv.cv_glm_mult_params(lazy, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
    for a, b in zip(single_lazy["full_cv_results"], single["full_cv_results"]):
        assert np.array_equal(a["cv_coefs"], b["cv_coefs"]) and np.array_equal(a["model"].coef_, b["model"].coef_)
    print("one GPU: lazy design == built design, bit-identical coefficients", flush=True)
for it in range(2):
    res = sglm_dist.cv_grid_strong(X0, shifts, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
    tl = sglm_dist.last_timeline
    print(f"rank {rank} pass {it} timeline ms: " + ", ".join(f"{k} {v:.1f}" if isinstance(v, float) else f"{k} {v}" for k, v in tl.items()), flush=True)
if rank == 0:
    assert res["best_params"] == single["best_params"], (res["best_params"], single["best_params"])
    assert abs(res["best_score"] - single["best_score"]) < 1e-12
    worst = 0.0
    for a, b in zip(res["full_cv_results"], single["full_cv_results"]):
        assert a["glm_kwargs"] == b["glm_kwargs"]
        # same statistics bits -> identical iterates: coefficients and intercepts are compared with ==; the scores are
        # quadratic forms whose row split depends on how many models a rank scores at once (last-bit differences)
        for k in ("cv_coefs", "cv_intercepts"):
            assert np.array_equal(a[k], b[k]), (a["glm_kwargs"], k, float(np.abs(a[k] - b[k]).max()))
        for k in ("cv_scores_train", "cv_scores_test"):
            assert np.allclose(a[k], b[k], rtol=0, atol=1e-12), (a["glm_kwargs"], k, float(np.abs(a[k] - b[k]).max()))
        assert np.array_equal(a["model"].coef_, b["model"].coef_) and a["model"].intercept_ == b["model"].intercept_
        assert abs(a["cv_R2_score"] - b["cv_R2_score"]) < 1e-12 and abs(a["cv_mse_score"] - b["cv_mse_score"]) < 1e-12
    print(f"STRONG CHECK OK: {world} GPUs == 1 GPU, {len(grid)} parameter sets x 6 fits, bit-identical", flush=True)
dist.barrier()
dist.destroy_process_group()
