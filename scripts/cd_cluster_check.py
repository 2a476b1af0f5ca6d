"""GPU check of the cluster coordinate-descent kernel against the first-generation kernel:
identical iterates (W bit-equal, same n_iter / update counts) and timing per (group, cluster) shape."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_pp

T, P = int(os.environ.get("DIAG_T", 200000)), int(os.environ.get("DIAG_P", 40))
lo, hi = int(os.environ.get("DIAG_LO", -20)), int(os.environ.get("DIAG_HI", 29))
NA = int(os.environ.get("DIAG_ALPHAS", 50))
shifts = [0] + [s for s in range(lo, hi + 1) if s != 0]
X0 = synth_data.synth_base(T, P, 1234)
beta = synth_data.synth_kernels(P, shifts, 1234)
d = sglm_pp.timeshift_multiple(torch.from_numpy(X0).cuda(), shift_amt_list=shifts)[hi:T + lo]
y = d @ torch.from_numpy(beta).cuda()
y = y + torch.randn_like(y) * y.std() * 1.5
y = (y - y.mean()) / y.std()
n, C = d.shape
folds = synth_data.synth_folds(n, 5, 1234)
W = torch.stack([torch.ones(n, dtype=torch.float64, device="cuda")] + [eng.index_counts(b, n) for _, b in folds])
G = eng.suffstats(d, y[:, None].contiguous(), W, [n] + [len(b) for _, b in folds])
probs = [eng.center(G[0], None, C, 1, 0, True)] + [eng.center(G[0], G[1 + f], C, 1, 0, True) for f in range(5)]
eng.fetch_scalars(probs)
alphas = np.logspace(-4, 0, NA); l1s = np.linspace(0.1, 0.9, 5)
ms = [eng.ModelSpec(p, "enet", a, l, 1000, 1e-4) for l in l1s for a in alphas for p in probs]
print(f"n={n} C={C} models={len(ms)}", flush=True)

def run(g, k, label, models=ms, plan=None):
    eng.CD_GROUP, eng.CD_CLUSTER, eng.CD_PLAN = (g, k, None) if plan is None else (None, None, plan)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Wd, info, st = eng.solve_models(models, C)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    h = int(np.argmax(info[:, 3]))
    print(f"{label:22s} time={dt*1e3:8.1f} ms updates={info[:,3].sum():.4e} alg GB/s={info[:,3].sum()*8*C/dt/1e9:8.1f} "
          f"max n_iter={info[:,2].max():.0f} p1share(heaviest)={info[h,5]:.3f}", flush=True)
    return Wd.cpu().numpy(), info

shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("DIAG_SHAPES", "1x1,1x4,2x4,4x4").split(",")]
W0, i0 = run(0, 0, "gen1 (warm-up)")
W0, i0 = run(0, 0, "gen1")
for g, k in shapes:
    try:
        run(g, k, f"cluster M={g} K={k} (wu)")
        W1, i1 = run(g, k, f"cluster M={g} K={k}")
    except Exception as e:
        print(f"cluster M={g} K={k}: FAILED {e}", flush=True)
        continue
    same = np.array_equal(W0, W1)
    dn = int(np.sum(i0[:, 2] != i1[:, 2])); du = int(np.sum(i0[:, 3] != i1[:, 3]))
    err = float(np.max(np.abs(W0 - W1)) / max(np.max(np.abs(W0)), 1e-300))
    gap_err = float(np.max(np.abs(i0[:, 0] - i1[:, 0]) / np.maximum(np.abs(i0[:, 0]), 1e-300)))
    print(f"    W bit-equal={same} max rel dW={err:.3e} models with different n_iter={dn} different updates={du} gap rel diff={gap_err:.2e}", flush=True)
for plan in [t for t in os.environ.get("DIAG_PLANS", "").split(";") if t]:
    run(0, 0, f"plan {plan} (wu)", plan=plan)
    W1, i1 = run(0, 0, f"plan {plan}", plan=plan)
    print(f"    W bit-equal={np.array_equal(W0, W1)} different n_iter={int(np.sum(i0[:, 2] != i1[:, 2]))}", flush=True)
if os.environ.get("DIAG_HEAVY"):
    order = np.argsort(-i0[:, 3])
    hv = ms[order[0]]
    peers = [m_ for m_ in ms if m_.problem is hv.problem and m_.l1_ratio == hv.l1_ratio]
    peers.sort(key=lambda m_: m_.alpha)
    run(0, 0, "heaviest alone gen1", [hv])
    for g, k in shapes:
        for sel in range(6):
            eng.CD_DEBUG_TIMER = sel
            run(g, k, f"heaviest alone M={g} K={k} timer{sel}", [hv])
        eng.CD_DEBUG_TIMER = 0
        if g > 1:
            run(g, k, f"heaviest {g} of a path", peers[:g])
