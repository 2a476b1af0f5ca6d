"""Timeline of BASELINE config 2 (Ridge CV grid, 5 folds x 20 alphas, 500k x 1220)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import synth_data, sglm_pp, sglm_cv, _sglm_native as nat
T, P = 500_000, 20
shifts = [0] + [s for s in range(-30, 31) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 2)).cuda()
beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 2)).cuda()
X = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[30:T - 30]
torch.manual_seed(2)
s = X @ beta; y = s + 1.5 * s.std() * torch.randn_like(s); y = ((y - y.mean()) / y.std()).contiguous()
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(X.shape[0], 5, 2)]
grid = [dict(alpha=float(a), l1_ratio=0, max_iter=1000, fit_intercept=True) for a in np.logspace(-3, 3, 20)]
def step():
    return sglm_cv.cv_glm_mult_params(X, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
step(); step(); torch.cuda.synchronize()
nat.enable_timing(True); nat.collect_timing()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
nat.collect_timing()
print(f"step {e0.elapsed_time(e1):.1f} ms")
prev = 0.0
for name, a, b in sorted(nat.last_intervals, key=lambda t: t[1]):
    if b - a > 0.05 or a - prev > 0.3:
        print(f"{a:8.2f} -> {b:8.2f}  ({b - a:7.2f} ms, gap before {a - prev:6.2f})  {name}")
    prev = max(prev, b)
