"""
bench.py — CV model fits/sec of the sGLM hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                   (CPU arm: the reference's own path)

One "step" = one pass of the hot path over one synthetic session:
    lag/shift gather  (T x P base signals  ->  T x C design, NaN edge rows dropped)
 -> CV grid of penalised Gaussian GLM fits: n_sets x (n_folds + 1) fits, the count the
    reference performs in cv_glm_mult_params (backend/sglm_cv.py:210-428).
Workload (default) = BASELINE.json configs[2], the configuration the north-star target is
quoted on and the largest that fits one GPU: ElasticNet grid 5 folds x 50 alphas x 5
l1_ratios on a 2M-timepoint x 2000-column lagged design (P=40 base signals x 50 shifts).
`value` is device-resident throughput (base signals, response and fold indices already in
HBM); `e2e` goes through the public drop-in API from pinned HOST buffers, host->device
copies and the device->host read of every result inside the timed region.
Multi-GPU: the path shards by independent sessions (configs[4]); each rank runs its own
session, no data-path collective, weak scaling; timing = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sabatinilab-glm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth_data  # noqa: E402

METRIC = "cv_model_fits_per_sec"
UNIT = "fits/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--T", type=int, default=2_000_000)
    ap.add_argument("--P", type=int, default=40)
    ap.add_argument("--shift-lo", type=int, default=-20)
    ap.add_argument("--shift-hi", type=int, default=29)
    ap.add_argument("--folds", type=int, default=5)
    ap.add_argument("--alphas", type=int, default=50)
    ap.add_argument("--l1s", type=int, default=5)
    ap.add_argument("--max-iter", type=int, default=1000)
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--warm-path", action="store_true",
                    help="opt-in non-reference mode: warm-started alpha paths (NOT the parity configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-T", type=int, default=40_000)
    return ap.parse_args()


def workload(args):
    shifts = [0] + [s for s in range(args.shift_lo, args.shift_hi + 1) if s != 0]
    alphas = np.logspace(-4, 0, args.alphas)
    l1s = np.linspace(0.1, 0.9, args.l1s) if args.l1s > 1 else np.array([0.5])
    grid = [dict(alpha=float(a), l1_ratio=float(round(l, 6)), max_iter=args.max_iter, fit_intercept=True,
                 tol=args.tol) for l in l1s for a in alphas]
    return shifts, grid


def make_session(args, seed, T):
    """Host-side synthetic session: base signals, kernels, noise, folds (design built later)."""
    shifts, _ = workload(args)
    X0 = synth_data.synth_base(T, args.P, seed)
    beta = synth_data.synth_kernels(args.P, shifts, seed)
    h_lo, h_hi = max(0, max(shifts)), max(0, -min(shifts))     # NaN rows at the top / bottom
    n_valid = T - h_lo - h_hi
    folds = synth_data.synth_folds(n_valid, args.folds, seed, group=1000)
    return X0, beta, (h_lo, h_hi), folds


def config_dict(args, n_gpus):
    shifts, grid = workload(args)
    C = args.P * len(shifts)
    return {"workload": f"ElasticNet CV grid {args.folds} folds x {args.alphas} alphas x {args.l1s} l1_ratios "
                        f"(+1 full-data refit per set = {len(grid) * (args.folds + 1)} fits/step) on "
                        f"{args.T} timepoints x {C} lagged columns (P={args.P} base signals x {len(shifts)} shifts), "
                        f"fp64, tol={args.tol}, max_iter={args.max_iter}, "
                        f"{'WARM-STARTED alpha paths (non-reference mode)' if args.warm_path else 'cold start'}, cyclic CD "
                        f"(BASELINE.json configs[2]); one independent session per GPU",
            "T": args.T, "C": C, "folds": args.folds, "alphas": args.alphas, "l1_ratios": args.l1s,
            "fits_per_step": len(grid) * (args.folds + 1), "sessions": n_gpus,
            "l2_policy": "inputs larger than L2 (design matrix %.1f GB >> 126 MB)" % (args.T * C * 8 / 1e9)}


# --------------------------------------------------------------------------- #
# clocks sampling during the timed region
# --------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- #
# CPU arm: the reference's path on host cores (oracle port, scikit-learn numerics)
# --------------------------------------------------------------------------- #
def cpu_reference_run(args, steps, warmup):
    """Times the reference's CPU path — oracle restatement of backend/sglm_pp.py +
    backend/sglm_cv.py + backend/sglm.py handing each fit to the installed scikit-learn
    (exactly what the reference executes), 4 fold threads as in backend/sglm_cv.py:162-170,
    BLAS/OpenMP at their defaults (all host cores) — on a BOUNDED sample of the workload."""
    from oracle import sglm_oracle as orc
    import warnings
    warnings.filterwarnings("ignore")
    try:    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=len(os.sched_getaffinity(0)))
    except Exception:
        pass
    shifts, grid = workload(args)
    Ts = min(args.cpu_sample_T, args.T)
    X0 = synth_data.synth_base(Ts, args.P, 1234)
    beta = synth_data.synth_kernels(args.P, shifts, 1234)
    sub = [dict(grid[i]) for i in (len(grid) // 2, len(grid) // 2 + args.alphas // 2, len(grid) - 1)][:3]
    n_folds = 2
    times, fits = [], len(sub) * (n_folds + 1)
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        Xd = orc.timeshift_multiple(X0, shift_amt_list=shifts)
        Xd = Xd[~np.isnan(Xd).any(axis=1)]
        if it == 0:
            y = synth_data.synth_response(Xd, beta, 1234)
            folds = synth_data.synth_folds(Xd.shape[0], n_folds, 1234, group=1000)
        orc.cv_glm_mult_params(Xd, y, folds, "Gaussian", [dict(g) for g in sub], score_method="r2",
                               engine="sklearn", n_threads=4)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    C = args.P * len(shifts)
    sample = (f"{len(sub)} param sets x ({n_folds} folds + refit) = {fits} sklearn ElasticNet fits on "
              f"{Ts} timepoints x {C} columns (T/{args.T // Ts} of the workload; per-fit cost is ~linear in T), "
              f"incl. the host lag gather; scikit-learn {__import__('sklearn').__version__}")
    return fits / sec, sec, sample, args.T / Ts


# --------------------------------------------------------------------------- #
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    cores = len(os.sched_getaffinity(0))

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
        v, sec, sample, scale = cpu_reference_run(args, steps, warmup)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, n_gpus),
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                 "extrapolated_full_size_value": v / scale},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import _sglm_native as nat
    import sglm_cv
    import sglm_pp
    import _engine
    _engine.WARM_START_PATHS = bool(args.warm_path)

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    shifts, grid = workload(args)
    C = args.P * len(shifts)
    fits_per_step = len(grid) * (args.folds + 1)

    # ---- one synthetic session per rank (host), then device-resident copies
    X0_h, beta, (h_lo, h_hi), folds_h = make_session(args, 1234 + rank, args.T)
    X0_pin = torch.from_numpy(X0_h).pin_memory()
    X0_d = X0_pin.to("cuda")
    design = sglm_pp.timeshift_multiple(X0_d, shift_amt_list=shifts)
    Xv = design[h_lo: args.T - h_hi]
    noise_seed = 99 + rank
    s = Xv @ torch.from_numpy(beta).cuda()                      # data generation only (not timed)
    from scipy.signal import lfilter
    e = lfilter([1.0], [1.0, -0.95], np.random.default_rng(noise_seed).standard_normal(Xv.shape[0]))
    e = torch.from_numpy(e / e.std()).cuda() * s.std() * float(np.sqrt(0.7 / 0.3))
    y_d = s + e
    y_d = ((y_d - y_d.mean()) / y_d.std()).contiguous()
    y_pin = y_d.cpu().pin_memory()
    del s, e, design, Xv
    folds_pin = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in folds_h]
    folds_d = [(a.to("cuda"), b.to("cuda")) for a, b in folds_pin]
    torch.cuda.synchronize()

    def step_device():
        d = sglm_pp.timeshift_multiple(X0_d, shift_amt_list=shifts)
        return sglm_cv.cv_glm_mult_params(d[h_lo: args.T - h_hi], y_d, folds_d, "Gaussian", [dict(g) for g in grid],
                                          score_method="r2")

    def step_e2e():
        x0 = X0_pin.to("cuda", non_blocking=True)
        yy = y_pin.to("cuda", non_blocking=True)
        fd = [(a.to("cuda", non_blocking=True), b.to("cuda", non_blocking=True)) for a, b in folds_pin]
        d = sglm_pp.timeshift_multiple(x0, shift_amt_list=shifts)
        return sglm_cv.cv_glm_mult_params(d[h_lo: args.T - h_hi], yy, fd, "Gaussian", [dict(g) for g in grid],
                                          score_method="r2")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    for _ in range(args.warmup):
        res = step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = nat.launches()
    nat.enable_timing(True)
    nat.collect_timing()
    total_ms, res = timed(step_device, args.steps)
    per_entry = nat.collect_timing()
    nat.enable_timing(False)
    launches = nat.launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = fits_per_step * args.steps * world / (total_ms / 1e3)

    # ---- per-kernel accounting for the roofline of the dominant kernel
    full = res["full_cv_results"]
    n_upd = float(sum(np.sum(r["_fit_info"]["cd_info"][:, 3]) for r in full))
    n_sweeps = float(sum(np.sum(r["_fit_info"]["cd_info"][:, 2]) for r in full))
    n_unconv = int(sum(np.sum(r["_fit_info"]["status"] != 0) for r in full))
    peaks = dict(FALLBACK_PEAKS)
    peaks_src = "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks.update(json.load(f))
            peaks_src = "measured"
    except Exception:
        pass
    tc_plan = nat.last_tc_plan
    n_rows = args.T - h_lo - h_hi
    n_test = sum(len(b) for _, b in folds_h)
    n_aug = C + 2
    kernels = {}
    for name, (calls, ms) in per_entry.items():
        kernels[name] = {"calls": calls, "ms_per_step": ms / args.steps}
    k_ms = {k: v["ms_per_step"] for k, v in kernels.items()}
    # the coordinate-descent grid is one logical kernel: its parts (clusters of heavy models, one CTA per
    # light model) are launched on two streams and overlap, so the time they cover together is what counts
    cd_names = ("sglm_enet_cd_cluster_f64", "sglm_enet_cd_gram_f64")
    cd_parts = {n: k_ms.pop(n) for n in cd_names if n in k_ms}
    CD = "sglm_enet_cd (cluster + per-model parts, concurrent)" if len(cd_parts) > 1 else next(iter(cd_parts), cd_names[0])
    if cd_parts:
        k_ms[CD] = nat.union_ms(cd_names) / args.steps
    if "sglm_gram_tc_cells_f64" in k_ms:          # the Gram over the disjoint cells of the row sets (same statistics)
        k_ms["sglm_gram_tc_f64"] = k_ms.pop("sglm_gram_tc_cells_f64")
    dominant = max(k_ms, key=k_ms.get)
    alg = {
        "sglm_timeshift_f64_ranged": ("hbm", 8.0 * args.T * args.P + 8.0 * args.T * C),
        "sglm_suffstats_f64": ("tensor", float(n_rows + n_test) * (n_aug * (n_aug + 1.0))),
        "sglm_gram_tc_f64": ("tensor", float(n_rows + n_test) * (n_aug * (n_aug + 1.0))),
        CD: ("hbm", n_upd * 8.0 * C + n_sweeps * 8.0 * 5 * C),
        "sglm_quadform_f64": ("hbm", 0.0),
    }
    bound, work = alg.get(dominant, ("hbm", 0.0))
    sec_dom = k_ms[dominant] / 1e3
    if bound == "hbm":
        achieved, peak, runit = work / sec_dom / 1e9, peaks["hbm_gbs"], "GB/s"
    else:
        achieved, peak, runit = work / sec_dom / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s"
    traffic = None
    try:    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get("sglm_enet_cd" if dominant == CD else dominant, {})
        traffic = tr.get("dram_bytes_per_launch")
        if dominant == CD and tr.get("dram_bytes_per_row_update"):
            traffic = tr["dram_bytes_per_row_update"] * n_upd       # scaled to this launch's coordinate updates
    except Exception:
        pass
    roofline = {"kernel": dominant, "bound": bound, "achieved": achieved, "peak": peak, "unit": runit,
                "frac": achieved / peak, "traffic": traffic, "peak_source": peaks_src,
                "note": ("algorithmic bytes = rows of Q (8*C bytes) per coordinate update that moved w + per-sweep "
                         "vectors; most row reads hit the 126 MB L2 (6 Gram matrices of 32 MB) and the cluster kernel "
                         "loads a moved row once for the 4 models of a group, so DRAM traffic is far below the "
                         "algorithmic bytes and frac can exceed 1 against the HBM copy peak"),
                "share_of_step": k_ms[dominant] / (total_ms / args.steps),
                "per_entry_ms_per_step": k_ms,
                "other": {
                    "gather_GBps": alg["sglm_timeshift_f64_ranged"][1] / (k_ms.get("sglm_timeshift_f64_ranged", np.inf) / 1e3) / 1e9,
                    "suffstats_fp64_TFLOPs_syrk_honest": alg["sglm_suffstats_f64"][1] / (k_ms.get("sglm_suffstats_f64", np.inf) / 1e3) / 1e12,
                    "gram_tc_useful_TFLOPs_syrk_honest": alg["sglm_gram_tc_f64"][1] / (k_ms.get("sglm_gram_tc_f64", np.inf) / 1e3) / 1e12,
                    "gram_tc_issued_int8_TOPs": (2.0 * tc_plan["tiles"] * 256 * 256 * tc_plan["n_pos"] / (k_ms.get("sglm_gram_tc_f64", np.inf) / 1e3) / 1e12) if tc_plan else None,
                    "gram_tc_plan": tc_plan,
                    "cd_GBps": alg[CD][1] / (k_ms.get(CD, np.inf) / 1e3) / 1e9,
                    "cd_parts_ms_per_step": cd_parts, "cd_plan": _engine._cd_plan(C, fits_per_step),
                    "cd_row_updates_per_step": n_upd,
                    "cd_sweeps_total": n_sweeps, "models_not_converged": n_unconv}}

    # ---- end to end through the public API from pinned host buffers
    e2e = None
    if not args.no_e2e:
        step_e2e()
        e2e_steps = max(1, min(args.steps, 2))
        e2e_ms, res_e = timed(step_e2e, e2e_steps)
        h2d = X0_pin.numel() * 8 + y_pin.numel() * 8 + sum((a.numel() + b.numel()) * 8 for a, b in folds_pin)
        d2h = sum(r["cv_coefs"].nbytes + r["cv_intercepts"].nbytes + r["cv_scores_train"].nbytes
                  + r["cv_scores_test"].nbytes + r["model"].coef_.nbytes + 8 * 6 for r in res_e["full_cv_results"])
        e2e = {"value": fits_per_step * e2e_steps * world / (e2e_ms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, sample, scale = cpu_reference_run(args, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "extrapolated_full_size_value": v / scale}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu,
                "best_params": res["best_params"], "best_score": float(res["best_score"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
