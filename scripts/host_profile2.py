"""cProfile of the host side of one full-size step (1 GPU): where the ~20 ms outside the kernels go."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm_cv, sglm_pp, sglm_dist
import torch.distributed as dist
T, P = 2_000_000, 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 1234)).cuda()
d = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[29:T - 20]
beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 1234)).cuda()
s = d @ beta
torch.manual_seed(0)
y = s + 1.5 * s.std() * torch.randn_like(s); y = ((y - y.mean()) / y.std()).contiguous()
del d, s
n = T - 49
folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(n, 5, 1234, group=1000)]
grid = [dict(alpha=float(a), l1_ratio=float(round(l, 6)), max_iter=1000, fit_intercept=True, tol=1e-4)
        for l in np.linspace(0.1, 0.9, 5) for a in np.logspace(-4, 0, 50)]
def step():
    dd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)
    return sglm_cv.cv_glm_mult_params(dd[29:T - 20], y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
for _ in range(2): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable(); print("step s:", (time.perf_counter() - t0) / 3)
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(45); print(st.getvalue()[:9000])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
def step2():
    return sglm_dist.cv_grid_strong(X0, shifts, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2", rows=(29, T - 20))
for _ in range(2): step2()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step2()
torch.cuda.synchronize(); pr.disable(); print("strong(1) timeline:", sglm_dist.last_timeline)
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(40); print(st.getvalue()[:8000])
dist.destroy_process_group()
