"""Diagnostic (GPU): per-model work distribution and latency/bandwidth regime of the CD kernel."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_pp

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
def specs(max_iter=1000):
    return [eng.ModelSpec(p, "enet", a, l, max_iter, 1e-4) for l in l1s for a in alphas for p in probs]
def run(models, label):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Wd, info, st = eng.solve_models(models, C)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    upd, fetch = info[:, 3].sum(), info[:, 3].sum()
    print(f"{label:28s} models={len(models):5d} time={dt*1e3:9.1f} ms  updates={upd:.3e} fetched={fetch:.3e} "
          f"alg GB/s={upd*8*C/dt/1e9:8.1f} fetched GB/s={fetch*8*C/dt/1e9:8.1f} max n_iter={info[:,2].max():.0f} "
          f"max upd={info[:,3].max():.3e} p1share(heaviest)={info[np.argmax(info[:,3]),5]:.3f} blocks(heaviest)={info[np.argmax(info[:,3]),4]:.0f}")
    return info
ms = specs()
run(ms, "warm-up all")
info = run(ms, "all 1500")
order = np.argsort(-info[:, 3])
if os.environ.get("DIAG_FULL"):
    print("top-10 updates:", info[order[:10], 3], "n_iter:", info[order[:10], 2])
    print("quantiles of updates:", np.quantile(info[:, 3], [0.5, 0.9, 0.99, 1.0]))
run([ms[order[0]]], "heaviest alone")
run([ms[i] for i in order[:148]], "top 148")
run([ms[i] for i in order[296:]], "all but top 296")
