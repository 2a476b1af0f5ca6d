"""NCCL check of the sharded CV grid (run under torchrun with 2+ GPUs): broadcast the inputs from rank 0,
shard the parameter sets over the ranks, all_gather the records — the result must equal the single-GPU grid."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch, torch.distributed as dist
import synth_data, sglm_pp, sglm_cv, sglm_dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
T, P = 60_000, 26
shifts = [0] + [s for s in range(-20, 30) if s != 0]            # 1300 columns: the wide-design plan on every rank
if rank == 0:
    X0 = synth_data.synth_base(T, P, 77)
    Xd = sglm_pp.timeshift_multiple(torch.from_numpy(X0).cuda(), shift_amt_list=shifts)[29:T - 20]
    y = synth_data.synth_response(Xd.cpu().numpy(), synth_data.synth_kernels(P, shifts, 77), 77)
    folds = synth_data.synth_folds(Xd.shape[0], 3, 77, group=500)
    Xb, yb, fb = sglm_dist.broadcast_inputs(Xd, y, folds, src=0)
else:
    Xb, yb, fb = sglm_dist.broadcast_inputs(None, None, None, src=0)
grid = [dict(alpha=float(a), l1_ratio=float(l), max_iter=1000, fit_intercept=True) for l in (0.2, 0.8) for a in np.logspace(-3, 0, 12)]
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
res = sglm_dist.cv_glm_mult_params_sharded(Xb, yb, fb, "Gaussian", [dict(g) for g in grid], score_method="r2")
torch.cuda.synchronize(); dist.barrier(); t_sh = time.perf_counter() - t0
if rank == 0:
    t0 = time.perf_counter()
    ref = sglm_cv.cv_glm_mult_params(Xb, yb, fb, "Gaussian", [dict(g) for g in grid], score_method="r2")
    torch.cuda.synchronize(); t_one = time.perf_counter() - t0
    assert res["best_params"] == ref["best_params"], (res["best_params"], ref["best_params"])
    worst = 0.0
    for a, b in zip(res["full_cv_results"], ref["full_cv_results"]):
        assert np.array_equal(a["cv_coefs"], b["cv_coefs"]) and np.array_equal(a["model"].coef_, b["model"].coef_)
        worst = max(worst, float(np.max(np.abs(a["cv_scores_test"] - b["cv_scores_test"]))))
    print(f"dist_check ok: world={world}, {len(grid)} parameter sets x 4 fits, coefficients identical, "
          f"score diff {worst:.1e}, sharded {t_sh*1e3:.1f} ms vs one GPU {t_one*1e3:.1f} ms, best {res['best_params']}")
dist.barrier()
dist.destroy_process_group()
