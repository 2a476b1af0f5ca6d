"""Does the cluster CD kernel speed up when the Gram matrices it streams fit in L2?  Same 750 models either as every 2nd
model of the 6-problem grid (6 x 32 MB of Q) or as all models of 3 problems (3 x 32 MB)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_pp
torch.manual_seed(0)
T, P = 200000, 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
X0 = synth_data.synth_base(T, P, 1234)
beta = synth_data.synth_kernels(P, shifts, 1234)
d = sglm_pp.timeshift_multiple(torch.from_numpy(X0).cuda(), shift_amt_list=shifts)[29:T - 20]
y = d @ torch.from_numpy(beta).cuda()
y = y + torch.randn_like(y) * y.std() * 1.5
y = (y - y.mean()) / y.std()
n, C = d.shape
folds = synth_data.synth_folds(n, 5, 1234)
W = torch.stack([torch.ones(n, dtype=torch.float64, device="cuda")] + [eng.index_counts(b, n) for _, b in folds])
G = eng.suffstats(d, y[:, None].contiguous(), W, [n] + [len(b) for _, b in folds])
probs = [eng.center(G[0], None, C, 1, 0, True)] + [eng.center(G[0], G[1 + f], C, 1, 0, True) for f in range(5)]
eng.fetch_scalars(probs)
alphas = np.logspace(-4, 0, 50); l1s = np.linspace(0.1, 0.9, 5)
def models(pp, every=1):
    ms = [eng.ModelSpec(p, "enet", a, l, 1000, 1e-4) for l in l1s for a in alphas for p in pp]
    return ms[::every]
def run(ms, plan):
    eng.CD_PLAN = plan
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Wd, info, st = eng.solve_models(ms, C)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return dt * 1e3, info[:, 4].max(), info[:, 4].sum()
for plan in ["4x2@0.3,0x0", "4x2", "4x4@0.3,0x0"]:
    for name, ms in [("6 problems, every 2nd model", models(probs, 2)), ("3 problems, all models", models(probs[:3])),
                     ("2 problems, all models", models(probs[:2])), ("1 problem", models(probs[:1]))]:
        ms_t, mx, tot = run(ms, plan)
        print(f"plan {plan:14s} {name:30s} models {len(ms):4d}: {ms_t:7.1f} ms, heaviest {mx:.0f} blocks -> {ms_t * 1e3 / mx:.2f} us/block, total blocks {tot:.0f}", flush=True)
