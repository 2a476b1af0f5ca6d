"""Timers of the cluster CD kernel UNDER LOAD: the whole 1500-model grid with a given plan, debug timer `sel`
(0 register phase, 1 waits, 2 publish, 3 look-ahead, 4 top/prologue, 5 end-of-sweep) reported for the heaviest models."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_pp
torch.manual_seed(0)
T, P = int(os.environ.get("DIAG_T", 200000)), 40
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
ms = [eng.ModelSpec(p, "enet", a, l, 1000, 1e-4) for l in l1s for a in alphas for p in probs]
sub = int(os.environ.get("DIAG_SUBSET", 1))
def run(plan, sel, models):
    eng.CD_PLAN, eng.CD_DEBUG_TIMER = plan, sel
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Wd, info, st = eng.solve_models(models, C)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    h = np.argsort(-info[:, 4])[:3]
    return dt, info[h, 4], info[h, 5], info[:, 3].sum()
for item in os.environ.get("DIAG_PLANS", "4x2@0.3,0x0").split(";"):
    plan, _, var = item.partition("!")
    os.environ["SGLM_TUNING"] = "1"
    os.environ["SGLM_CDC_VARIANT"] = var or "0"
    for every in [int(v) for v in os.environ.get('DIAG_EVERY', '1,2,8').split(',')]:
        models = ms[::every] if every > 1 else ms
        run(plan, 0, models)
        out = []
        for sel in range(6):
            dt, blk, share, upd = run(plan, sel, models)
            out.append((dt, share[0]))
        names = ["register", "waits", "publish", "lookahead", "top", "end"]
        us = out[0][0] * 1e6 / blk[0]
        print(f"variant {var or '0'} plan {plan:24s} models {len(models):5d} time {out[0][0]*1e3:7.1f} ms heaviest blocks {blk[0]:.0f} -> {us:.2f} us/block; shares of the heaviest model: "
              + ", ".join(f"{nm} {s:.2f}" for nm, (_, s) in zip(names, out)), flush=True)
