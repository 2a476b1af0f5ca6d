"""Per-model work and timeline of the coordinate-descent launches (GPU): sweeps, coordinate updates, blocks walked,
start / end time (globaltimer) of every model under a given plan -> gpurun_out/cd_dump_<tag>.npz."""
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
meta = np.array([(a, l, pi) for l in l1s for a in alphas for pi in range(6)])
def run(plan, sel):
    eng.CD_PLAN, eng.CD_DEBUG_TIMER = plan, sel
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Wd, info, st = eng.solve_models(ms, C)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return dt, info
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for k, plan in enumerate(os.environ.get("DIAG_PLANS", "4x2@0.3,0x0").split(";")):
    run(plan, 0)
    dt, info = run(plan, 0)
    _, i6 = run(plan, 6)
    # start and end of one and the same launch cannot be had from one info column: two launches, each relative to its
    # own earliest start / to the same launch's starts is not possible -> report durations from a third column instead
    _, i7 = run(plan, 7)
    print(f"plan {plan}: {dt*1e3:.1f} ms  max blocks {info[:,4].max():.0f}  max sweeps {info[:,2].max():.0f}", flush=True)
    np.savez(os.path.join(ROOT, "gpurun_out", f"cd_dump_{os.environ.get('DIAG_TAG','a')}_{k}.npz"), meta=meta, info=info,
             start=i6[:, 5], dur=i7[:, 5], plan=plan, ms=dt * 1e3)
