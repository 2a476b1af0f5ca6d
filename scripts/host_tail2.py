"""Host wall time (no added synchronisation) of the tail of a full-size step: what runs after the last kernel."""
import os, sys, time, functools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm_cv, sglm_pp, _engine as eng
T, P = 2_000_000, 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 1234)).cuda()
n = T - 49
torch.manual_seed(0)
y = torch.randn(n, dtype=torch.float64, device="cuda")
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(n, 5, 1234, group=1000)]
grid = [dict(alpha=float(a), l1_ratio=float(round(l, 6)), max_iter=200, fit_intercept=True, tol=1e-4)
        for l in np.linspace(0.1, 0.9, 5) for a in np.logspace(-4, 0, 50)]
log = []
def wrap(obj, name):
    f = getattr(obj, name)
    @functools.wraps(f)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); log.append((name, t0, time.perf_counter())); return r
    setattr(obj, name, g)
for obj, name in [(sglm_cv.GaussianSession, "assemble"), (sglm_cv.GaussianSession, "download_coefficients"), (sglm_cv.GaussianSession, "download_and_assemble"),
                  (sglm_cv, "_select_best"), (sglm_cv, "_order_result"), (sglm_cv.GaussianSession, "score"), (sglm_cv.GaussianSession, "moments"),
                  (eng, "solve_models"), (sglm_cv, "_cv_batch"), (sglm_cv, "_gaussian_grid"), (sglm_cv, "cv_glm_mult_params"),
                  (sglm_pp.DeviceDesign, "dropna"), (sglm_pp, "timeshift_multiple"), (sglm_cv.GaussianSession, "build_statistics"),
                  (sglm_cv.GaussianSession, "model_specs"), (sglm_cv, "_normalise_cv_idx")]:
    wrap(obj, name)
def step():
    dd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, device=True).dropna()
    return sglm_cv.cv_glm_mult_params(dd, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
import gc
for _ in range(3): step()
gc.collect(); gc.freeze()
log.clear(); torch.cuda.synchronize(); t00 = time.perf_counter(); step(); t_ret = time.perf_counter(); torch.cuda.synchronize()
print(f"step returns after {(t_ret - t00) * 1e3:.2f} ms")
agg = {}
for name, a, b in log:
    c = agg.setdefault(name, [0, 0.0, None, None]); c[0] += 1; c[1] += (b - a) * 1e3
    c[2] = a if c[2] is None else c[2]; c[3] = b
for name, (cnt, ms, a, b) in sorted(agg.items(), key=lambda kv: kv[1][2]):
    print(f"  {name:26s} x{cnt:4d} {ms:8.2f} ms   first start {1e3 * (a - t00):8.2f}  last end {1e3 * (b - t00):8.2f}")
