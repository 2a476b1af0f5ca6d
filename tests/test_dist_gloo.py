"""world_size = 2 over gloo (CPU): broadcast of the inputs, round-robin sharding of the
parameter sets, all_gather of the per-set records, ordering and selection.  The per-rank
compute is the CPU oracle here (the B200 plan is what runs on a GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, ret):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sglm_dist
    from oracle import sglm_oracle as orc
    import synth_data

    X = y = cv_idx = None
    if rank == 0:
        X0 = synth_data.synth_base(900, 3, 5)
        shifts = [0, -1, 1]
        X = orc.timeshift_multiple(X0, shift_amt_list=shifts)[1:-1]
        y = synth_data.synth_response(X, synth_data.synth_kernels(3, shifts, 5), 5)
        cv_idx = synth_data.synth_folds(X.shape[0], 3, 5, group=100)
    Xt, yt, idx = sglm_dist.broadcast_inputs(X, y, cv_idx, src=0)
    Xn, yn = Xt.numpy(), yt.numpy()
    folds = [(a.numpy(), b.numpy()) for a, b in idx]

    def compute(X_, y_, cv_, entries, score_method):
        return [orc.cv_glm_single_params(Xn, yn, folds, name, kw, score_method) for name, kw in entries]

    class _M:
        def __init__(self, coef, icpt):
            self.coef_, self.intercept_ = coef, icpt

    grid = orc.generate_mult_params(dict(alpha=[1e-3, 1e-2, 1.0], l1_ratio=[0, 0.5, 1], roll=[0, 2]),
                                    dict(max_iter=500, fit_intercept=True))
    res = sglm_dist.cv_glm_mult_params_sharded(Xn, yn, folds, "Gaussian", [dict(g) for g in grid],
                                               score_method="r2", compute=compute,
                                               rebuild_model=lambda n, kw, c, b: _M(c, b))
    want = orc.cv_glm_mult_params(Xn, yn, folds, "Gaussian", [dict(g) for g in grid], score_method="r2")
    ok = res["best_params"] == want["best_params"] and abs(res["best_score"] - want["best_score"]) < 1e-12
    ok = ok and len(res["full_cv_results"]) == len(grid)
    for a, b in zip(res["full_cv_results"], want["full_cv_results"]):
        ok = ok and a["glm_kwargs"] == b["glm_kwargs"] and np.allclose(a["cv_coefs"], b["cv_coefs"], atol=1e-12)
        ok = ok and np.allclose(a["model"].coef_, b["model"].coef_, atol=1e-12)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cv_grid_world2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_strong(rank, world, port, ret):
    """Host logic of the strong-scaled grid (sglm_dist.cv_grid_strong) over gloo: models dealt by cost in a snake,
    every rank "solves" its share (the CPU oracle stands in for the B200 solvers), the packed records are gathered
    and re-ordered — the result must equal the single-process grid."""
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sglm_dist
    from oracle import sglm_oracle as orc
    rng = np.random.default_rng(3)                          # same data on both ranks
    n, C = 300, 7
    X = rng.standard_normal((n, C))
    y = X @ rng.standard_normal(C) + 0.3 * rng.standard_normal(n)
    alphas, l1s = [1e-3, 1e-2, 1e-1, 1.0], [0.2, 0.9]
    specs = [(a, l) for l in l1s for a in alphas] + [(0.5, 0.0)]

    class _S:
        def __init__(self, a, l):
            self.alpha, self.l1_ratio, self.kind = a, l, ("ridge" if l == 0.0 else "enet")
    costs = [sglm_dist.model_cost(_S(a, l)) for a, l in specs]
    owner, order = sglm_dist.deal_models(costs, world)
    counts = np.bincount(owner, minlength=world)
    ok = counts.max() - counts.min() <= 1 and costs[order[0]] == max(costs) and owner[order[0]] != owner[order[1]]
    lists = [np.flatnonzero(owner == r) for r in range(world)]
    mine = lists[rank]
    W = np.zeros((len(mine), C)); b = np.zeros(len(mine))
    for k, i in enumerate(mine):
        a, l = specs[i]
        w, b0 = (orc.ridge_fit(X, y, alpha=a) if l == 0.0 else orc.enet_fit(X, y, alpha=a, l1_ratio=l)[:2])
        W[k], b[k] = w, b0
    n_pad = max(len(o) for o in lists)
    z = torch.zeros(len(mine), dtype=torch.float64)
    pack = sglm_dist.pack_results(torch.from_numpy(W), torch.from_numpy(b), z, z + 1, z + 2, np.full((len(mine), 6), 7.0),
                                  np.asarray(mine), n_pad)
    parts = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(parts, pack)
    full = sglm_dist.unpack_results(torch.stack(parts), lists, C).numpy()
    for i, (a, l) in enumerate(specs):
        w, b0 = (orc.ridge_fit(X, y, alpha=a) if l == 0.0 else orc.enet_fit(X, y, alpha=a, l1_ratio=l)[:2])
        ok = ok and np.array_equal(full[i, :C], w) and full[i, C] == b0 and full[i, C + 10] == i
        ok = ok and full[i, C + 2] == 1.0 and full[i, C + 3] == 2.0 and np.all(full[i, C + 4:C + 10] == 7.0)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_strong_grid_host_logic_world2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_strong, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
