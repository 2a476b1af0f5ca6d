"""
sglm_dist — multi-GPU sharding of the CV grid (one process per GPU, torch.distributed).

The reference has no distributed backend (single process, Python threads).  The grid shards
naturally: every parameter set — its F fold fits plus the full-data refit — is independent
of every other (backend/sglm_cv.py:131, :308 already treats them as separate queue items).
So parameter sets are dealt round-robin (in cost order) to the ranks, every rank builds
the statistics of the shared design locally, and the only collectives are

    broadcast  : the inputs from rank `src` (base signals / design, response, fold indices)
    all_gather : the per-set result records (coefficients, intercepts, scores)

— no collective inside the data path.  NCCL carries the tensors when the process group is
NCCL (NVLink 5 / NVSwitch on a B200 box); the same code runs over gloo on CPU tensors,
which is how tests/test_dist_gloo.py exercises it with world_size = 2.

`compute` is the per-rank engine: by default the B200 batched plan of sglm_cv; tests inject
the CPU oracle so that the sharding / ordering / selection logic is checked without a GPU.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def shard_indices(n_items, world_size, rank, cost=None):
    """Indices of the items owned by `rank`: items sorted by decreasing cost and dealt
    round-robin, so every rank gets a similar mix of expensive (weakly regularised) and
    cheap parameter sets."""
    order = np.arange(n_items) if cost is None else np.argsort(-np.asarray(cost, dtype=np.float64), kind="stable")
    return sorted(int(i) for i in order[rank::world_size])


def default_cost(glm_kwargs):
    """Heuristic work estimate of one parameter set: coordinate descent on a weak penalty
    needs many sweeps; direct solvers (alpha == 0 or l1_ratio == 0) are cheap."""
    a = float(glm_kwargs.get("alpha", 1.0))
    l1 = float(glm_kwargs.get("l1_ratio", 0.5))
    if a == 0.0 or l1 == 0.0:
        return 0.0
    return 1.0 / max(a * max(l1, 1e-3), 1e-12)


def broadcast_inputs(X, y, cv_idx, src=0, group=None, device=None):
    """Broadcast (X, y, cv_idx) from rank `src`.  Non-source ranks pass None.  Tensors travel
    over the group's backend (NCCL: device tensors; gloo: CPU tensors)."""
    import torch
    dist = _dist()
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        Xt = torch.as_tensor(X, dtype=torch.float64)
        yt = torch.as_tensor(y, dtype=torch.float64).reshape(-1)
        meta = [dict(x_shape=tuple(Xt.shape), n_y=int(yt.numel()),
                     folds=[(int(len(a)), int(len(b))) for a, b in cv_idx])]
    dist.broadcast_object_list(meta, src=src, group=group)
    m = meta[0]
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    if rank == src:
        Xt, yt = Xt.to(dev).contiguous(), yt.to(dev).contiguous()
        idx = [(torch.as_tensor(np.asarray(a), dtype=torch.int64).to(dev),
                torch.as_tensor(np.asarray(b), dtype=torch.int64).to(dev)) for a, b in cv_idx]
    else:
        Xt = torch.empty(m["x_shape"], dtype=torch.float64, device=dev)
        yt = torch.empty(m["n_y"], dtype=torch.float64, device=dev)
        idx = [(torch.empty(na, dtype=torch.int64, device=dev), torch.empty(nb, dtype=torch.int64, device=dev))
               for na, nb in m["folds"]]
    dist.broadcast(Xt, src=src, group=group)
    dist.broadcast(yt, src=src, group=group)
    for a, b in idx:
        dist.broadcast(a, src=src, group=group)
        dist.broadcast(b, src=src, group=group)
    return Xt, yt, idx


def _b200_compute(X, y, cv_idx, entries, score_method):
    import sglm_cv
    return sglm_cv._cv_batch(X, y, cv_idx, entries, None, None, score_method)


def _record(r):
    """Picklable, device-free record of one result dict (the fitted model travels as its
    coefficients; it is rebuilt on the receiving side)."""
    m = r["model"]
    return {k: r[k] for k in ("cv_coefs", "cv_intercepts", "cv_scores_train", "cv_scores_test",
                              "cv_mean_score_train", "cv_mean_score", "cv_std_score", "cv_R2_score",
                              "cv_mse_score", "glm_kwargs")} | {
        "coef_": np.asarray(m.coef_), "intercept_": float(np.asarray(m.intercept_).reshape(-1)[0])}


def cv_glm_mult_params_sharded(X, y, cv_idx, model_name, glm_kwarg_lst, verbose=0, score_method='mse',
                               group=None, compute=None, rebuild_model=None):
    """Distributed `cv_glm_mult_params`: every rank calls it with the same arguments (use
    `broadcast_inputs` first when only one rank holds the data).  Returns, on every rank, the
    reference's result dict (backend/sglm_cv.py:420-426) with `full_cv_results` in the order
    of `glm_kwarg_lst` and the reference's strict-'>' selection."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    compute = compute or _b200_compute
    entries = []
    for kw in glm_kwarg_lst:
        name = kw.pop('model_name', 'Gaussian')                    # backend/sglm_cv.py:288
        entries.append((name, kw))
    mine = shard_indices(len(entries), world, rank, [default_cost(kw) for _, kw in entries])
    local = compute(X, y, cv_idx, [entries[i] for i in mine], score_method) if mine else []
    payload = [(i, _record(r)) for i, r in zip(mine, local)]
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    records = [None] * len(entries)
    for part in gathered:
        for i, rec in part:
            records[i] = rec
    for (name, kw), rec in zip(entries, records):                  # 'roll' was popped on the owning rank only
        kw.pop('roll', None)
        rec["glm_kwargs"] = kw
    resp = []
    for (name, kw), rec in zip(entries, records):
        out = dict(rec)
        coef, icpt = out.pop("coef_"), out.pop("intercept_")
        out["model"] = rebuild_model(name, kw, coef, icpt) if rebuild_model else _rebuild(name, kw, coef, icpt)
        resp.append(out)
    best_score, best = -np.inf, None
    for r in resp:
        s = r['cv_R2_score'] if score_method == 'r2' else r['cv_mean_score']
        if score_method in ('r2', 'mse') and s > best_score:
            best_score, best = s, r
    return {'best_score': best_score, 'best_score_std': best['cv_std_score'], 'best_params': best['glm_kwargs'],
            'best_model': best['model'], 'full_cv_results': resp}


def _rebuild(model_name, kw, coef, intercept):
    import sglm_
    glm = sglm_.GLM(model_name, **kw)
    glm.model.coef_ = np.asarray(coef, dtype=np.float64)
    glm.model.intercept_ = float(intercept)
    glm.coef_ = glm.beta_ = glm.model.coef_
    glm.intercept_ = glm.beta0_ = glm.model.intercept_
    return glm
