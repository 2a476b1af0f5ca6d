"""Host time of the parts of a full-size step that the GPU cannot hide (after the last kernel / before the first)."""
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
grid = [dict(alpha=float(a), l1_ratio=float(round(l, 6)), max_iter=30, fit_intercept=True, tol=1e-4)
        for l in np.linspace(0.1, 0.9, 5) for a in np.logspace(-4, 0, 50)]
log = []
def wrap(obj, name):
    f = getattr(obj, name)
    @functools.wraps(f)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); log.append((name, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(obj, name, g)
for obj, name in [(sglm_cv.GaussianSession, "assemble"), (sglm_cv.GaussianSession, "download_coefficients"), (sglm_cv, "_select_best"),
                  (sglm_cv, "_order_result"), (sglm_cv, "_normalise_cv_idx"), (sglm_cv.GaussianSession, "model_specs"),
                  (sglm_cv.GaussianSession, "score"), (sglm_cv.GaussianSession, "build_statistics"), (sglm_cv.GaussianSession, "__init__"),
                  (eng, "solve_models"), (sglm_pp.DeviceDesign, "dropna"), (sglm_pp, "timeshift_multiple"), (sglm_cv, "_cv_batch")]:
    wrap(obj, name)
def step():
    dd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, device=True).dropna()
    return sglm_cv.cv_glm_mult_params(dd, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
for _ in range(2): step()
log.clear(); torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize()
print(f"step (max_iter 30, every stage synchronised): {(time.perf_counter() - t0) * 1e3:.1f} ms")
from collections import OrderedDict
agg = OrderedDict()
for k, v in log: agg[k] = agg.get(k, 0.0) + v
for k, v in agg.items(): print(f"  {k:24s} {v:8.2f} ms")
