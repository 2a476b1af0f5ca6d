"""Timeline of one device-resident bench step: start/end of every ABI call (CUDA events on the launching
stream) and the gaps between them — where the GPU waits for the host."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm_pp, sglm_cv, _sglm_native as nat
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
    dd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, device=True).dropna()
    return sglm_cv.cv_glm_mult_params(dd, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
step(); step(); torch.cuda.synchronize()
nat.enable_timing(True); nat.collect_timing()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
nat.collect_timing()
print(f"step {e0.elapsed_time(e1):.1f} ms")
print("first call starts", f"{e0.elapsed_time(nat._first_event):.2f} ms after the step's start event" if getattr(nat, "_first_event", None) is not None else "")
prev = 0.0
for name, a, b in sorted(nat.last_intervals, key=lambda t: t[1]):
    print(f"{a:8.2f} -> {b:8.2f}  ({b - a:7.2f} ms, gap before {a - prev:6.2f})  {name}")
    prev = max(prev, b)
