"""torchrun --nproc-per-node N scripts/strong_profile.py : per-rank stage timeline of sglm_dist.cv_grid_strong at full size and
a host profile of rank 0 (where do the milliseconds outside the kernels go)."""
import cProfile, io, os, pstats, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sabatinilab-glm_b200")):
    sys.path.insert(0, p)
import synth_data, sglm_dist  # noqa: E402
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
T, P = 2_000_000, 40
shifts = [0] + [s for s in range(-20, 30) if s != 0]
grid = [dict(alpha=float(a), l1_ratio=float(l), max_iter=int(os.environ.get("MAX_ITER", 1000)), fit_intercept=True, tol=1e-4)
        for l in np.linspace(0.1, 0.9, 5) for a in np.logspace(-4, 0, 50)]
X0 = y = folds = None
if rank == 0:
    X0 = torch.from_numpy(synth_data.synth_base(T, P, 1234)).cuda()
    y = torch.randn(T - 49, dtype=torch.float64, device="cuda")
    folds = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in synth_data.synth_folds(T - 49, 5, 1234)]
def step():
    return sglm_dist.cv_grid_strong(X0, shifts, y, folds, "Gaussian", [dict(g) for g in grid], score_method="r2", rows=(29, T - 20))
for _ in range(2): step()
import gc; gc.collect(); gc.freeze()
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: " + ", ".join(f"{k} {v:.1f}" if isinstance(v, float) else f"{k} {v}" for k, v in sglm_dist.last_timeline.items()), flush=True)
if rank == 0:
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22); print(st.getvalue()[:5000], flush=True)
dist.barrier(); dist.destroy_process_group()
