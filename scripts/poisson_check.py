import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm_pp, _engine as eng
T, P = 1_000_000, 20
shifts = [0] + [s for s in range(-20, 20) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 4)).cuda()
beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 4)).cuda()
X = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[19:T - 20]
s = X @ beta; z = (s - s.mean()) / s.std()
y = torch.poisson(torch.exp(0.3 * z - 1.0)).contiguous()
print("shape", X.shape, "mean y", float(y.mean()))
for alpha in (1e-3, 1e-1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w, b, n_it = eng.poisson_irls(X, y, alpha, True, None, 100, 1e-4)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    eta = X @ torch.from_numpy(w).cuda() + b
    mu = torch.exp(eta)
    g = (X.T @ (mu - y)) / X.shape[0] + alpha * torch.from_numpy(w).cuda()
    print(f"alpha={alpha} n_iter={n_it} time={dt*1e3:.1f} ms  |grad|_inf={float(g.abs().max()):.3e}  |w|_inf={np.abs(w).max():.4f} b={b:.4f}")
