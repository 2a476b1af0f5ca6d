"""cProfile (by own time) of one full-size lazy-design step: which host functions cost milliseconds."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm_cv, sglm_pp
T, P = 2_000_000, 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 1234)).cuda()
n = T - 49
torch.manual_seed(0)
y = torch.randn(n, dtype=torch.float64, device="cuda")
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(n, 5, 1234, group=1000)]
grid = [dict(alpha=float(a), l1_ratio=float(round(l, 6)), max_iter=int(os.environ.get("MAX_ITER", 30)), fit_intercept=True, tol=1e-4)
        for l in np.linspace(0.1, 0.9, 5) for a in np.logspace(-4, 0, 50)]
def step():
    dd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, device=True).dropna()
    return sglm_cv.cv_glm_mult_params(dd, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
for _ in range(3): step()
import gc; gc.collect(); gc.freeze()
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable(); print("step ms:", (time.perf_counter() - t0) / 3 * 1e3)
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(40); print(st.getvalue()[:7000])
