"""Single heavy model through the CD kernel (for ncu source-level profiling)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_pp
T, P = 60000, 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
X0 = synth_data.synth_base(T, P, 1234)
beta = synth_data.synth_kernels(P, shifts, 1234)
d = sglm_pp.timeshift_multiple(torch.from_numpy(X0).cuda(), shift_amt_list=shifts)[29:T - 20]
y = d @ torch.from_numpy(beta).cuda()
y = y + torch.randn_like(y) * y.std() * 1.5
y = (y - y.mean()) / y.std()
n, C = d.shape
G = eng.suffstats(d, y[:, None].contiguous())
p = eng.center(G[0], None, C, 1, 0, True)
eng.fetch_scalars([p])
n_models = int(os.environ.get("N_MODELS", 1))
mi = int(os.environ.get("MAX_ITER", 60))
ms = [eng.ModelSpec(p, "enet", 1e-4, 0.1, mi, 1e-4) for _ in range(n_models)]
import time
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Wd, info, st = eng.solve_models(ms, C)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"models={n_models} max_iter={mi} time={dt*1e3:.1f} ms updates/model={info[0,3]:.0f} us/update={dt*1e6/info[0,3]:.3f}")
