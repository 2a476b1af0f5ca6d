"""Wall-clock of the other BASELINE.json configurations on one B200 (device-resident inputs)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm, sglm_pp, sglm_cv

def session(T, P, lo, hi, seed, poisson=False):
    shifts = [0] + [s for s in range(lo, hi + 1) if s != 0]
    X0 = torch.from_numpy(synth_data.synth_base(T, P, seed)).cuda()
    beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, seed)).cuda()
    h_lo, h_hi = max(0, hi), max(0, -lo)
    def design():
        return sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[h_lo:T - h_hi]
    X = design()
    torch.manual_seed(seed)           # reproducible noise: the selected alpha is part of the record
    s = X @ beta
    if poisson:
        z = (s - s.mean()) / s.std()
        y = torch.poisson(torch.exp(0.3 * z - 1.0))
    else:
        y = s + 1.5 * s.std() * torch.randn_like(s)
        y = (y - y.mean()) / y.std()
    return design, y.contiguous(), X.shape

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r

out = {}
# c1: single ElasticNet fit, 100k x 10 predictors x 41 shifts
design, y, shp = session(100_000, 10, -20, 20, 1)
sec, _ = timed(lambda: sglm.GLM("Gaussian", alpha=0.01, l1_ratio=0.5).fit(design(), y))
out["c1_single_enet_fit_100k_x_410"] = {"seconds": sec, "fits_per_s": 1 / sec, "shape": list(shp)}
# c2: Ridge CV grid 5 folds x 20 alphas, 500k x 20 x 61
design, y, shp = session(500_000, 20, -30, 30, 2)
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(shp[0], 5, 2)]
grid = [dict(alpha=float(a), l1_ratio=0, max_iter=1000, fit_intercept=True) for a in np.logspace(-3, 3, 20)]
sec, r = timed(lambda: sglm_cv.cv_glm_mult_params(design(), y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2"))
out["c2_ridge_cv_5x20_500k_x_1220"] = {"seconds": sec, "fits_per_s": 120 / sec, "shape": list(shp), "best": r["best_params"]}
# c4: Poisson alpha sweep (reduced: 3 alphas x (2 folds + refit)), 1M x 800
design, y, shp = session(1_000_000, 20, -20, 19, 4, poisson=True)
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(shp[0], 2, 4)]
grid = [dict(alpha=float(a), model_name="Poisson") for a in (1e-3, 1e-1, 10.0)]
X = design()
sec, r = timed(lambda: sglm_cv.cv_glm_mult_params(X, y, folds, "Poisson", [dict(g) for g in grid], score_method="r2"), reps=1)
out["c4_poisson_3alphas_x_(2folds+refit)_1M_x_800"] = {"seconds": sec, "fits_per_s": 9 / sec, "shape": list(shp), "best": r["best_params"]}
print(json.dumps(out, indent=1))
