"""cProfile of one device-resident bench step (host-side overhead of the launch plan)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm_pp, sglm_cv
T, P = int(os.environ.get("HP_T", 2_000_000)), 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 1234)).cuda()
beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 1234)).cuda()
d = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)
Xv = d[29:T - 20]
y = Xv @ beta; y = y + 1.5 * y.std() * torch.randn_like(y); y = ((y - y.mean()) / y.std()).contiguous()
del d, Xv
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(T - 49, 5, 1234)]
grid = [dict(alpha=float(a), l1_ratio=float(l), max_iter=1000, fit_intercept=True, tol=1e-4)
        for l in np.linspace(0.1, 0.9, 5) for a in np.logspace(-4, 0, 50)]
def step():
    dd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)
    r = sglm_cv.cv_glm_mult_params(dd[29:T - 20], y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
    torch.cuda.synchronize()
    return r
step(); step()
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
