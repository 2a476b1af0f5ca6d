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
