"""profiles/ncu_traffic.json from an `ncu --page raw --csv` export: per-launch DRAM bytes and unit utilisations of the hot
kernels (what bench.py quotes as `roofline.traffic`).  Usage: python scripts/ncu_traffic.py raw.csv out.json "<comment>" [bench.json [previous_ncu_traffic.json]] """
import csv, json, sys

WANT = {
    "gpu__time_duration.sum": ("duration_ms", None),
    "dram__bytes_read.sum": ("dram_read", None),
    "dram__bytes_write.sum": ("dram_write", None),
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": ("dram_throughput_pct", 1.0),
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": ("lts_throughput_pct", 1.0),
    "lts__t_sector_hit_rate.pct": ("l2_hit_rate_pct", 1.0),
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": ("tensor_pipe_active_pct", 1.0),
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active": ("tensor_imma_active_pct", 1.0),
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": ("tensor_cycles_active_pct", 1.0),
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": ("fp64_pipe_pct", 1.0),
    "smsp__issue_active.avg.pct_of_peak_sustained_active": ("issue_active_pct", 1.0),
    "sm__warps_active.avg.pct_of_peak_sustained_active": ("warps_active_pct", 1.0),
    "launch__registers_per_thread": ("registers_per_thread", 1.0),
    "launch__grid_size": ("grid", 1.0),
    "launch__cluster_size": ("cluster_size", 1.0),
}
UNIT = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(raw, out, comment):
    rows = list(csv.reader(open(raw, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    res = {"_comment": comment}
    seen = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kname = r[col["Kernel Name"]]
        short = kname.split("(")[0].split("<")[0].split("::")[-1].strip()
        k = seen.get(short, 0)
        seen[short] = k + 1
        key = short if k == 0 else f"{short}#{k + 1}"
        e = {"kernel": kname[:160]}
        for metric, (field, scale) in WANT.items():
            if metric not in col:
                continue
            v = num(r[col[metric]])
            if v is None:
                e[field] = None
                continue
            u = units[col[metric]]
            e[field] = v * (UNIT.get(u, 1.0) if scale is None else scale)
        if e.get("dram_read") is not None and e.get("dram_write") is not None:
            e["dram_bytes_per_launch"] = e.pop("dram_read") + e.pop("dram_write")
        res[key] = e
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res.items():
        if k != "_comment":
            print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "kernel"})


def finish(out, bench_json, previous_json):
    """Post-processing bench.py relies on: NaN -> null; counters of a kernel whose capture did not complete are taken from the
    previous capture (noted in the entry); `sglm_enet_cd` = DRAM bytes of all coordinate-descent launches of the plan (under ncu
    they are serialised) per step and per coordinate update; the gather's entry is carried over."""
    import math
    res = json.load(open(out))
    prev = json.load(open(previous_json)) if previous_json else {}
    bench = json.load(open(bench_json))

    def clean(o):
        if isinstance(o, dict):
            return {k: clean(v) for k, v in o.items()}
        return None if isinstance(o, float) and math.isnan(o) else o
    res = clean(res)
    solo_key = next((k for k in res if "enet_cd_gram_kernel" in k), None)
    if solo_key:
        solo = res.pop(solo_key)
        old = prev.get("enet_cd_gram_kernel", {})
        filled = [f for f in solo if solo[f] is None and old.get(f) is not None]
        for f in filled:
            solo[f] = old[f]
        if filled:
            solo["note"] = "counters that did not complete in this capture are those of the previous full-size capture of the same kernel and launch shape"
        res["enet_cd_gram_kernel"] = solo
    cd = [v for k, v in res.items() if k.startswith("enet_cd") and isinstance(v, dict)]
    tot = sum(v.get("dram_bytes_per_launch") or 0.0 for v in cd)
    upd = bench["roofline"]["coordinate_updates_per_step"]
    res["sglm_enet_cd"] = {"dram_bytes_per_launch": tot, "dram_bytes_per_row_update": tot / upd,
                           "note": "sum over the serialised coordinate-descent launches of the plan / coordinate updates of the step (%.0f); "
                                   "serialised launches hit L2 less often than the concurrent ones of the benchmark (the (4,4) head alone: 5 %%), "
                                   "so this is an upper bound" % upd}
    if "sglm_timeshift_f64_ranged" in prev:
        res["sglm_timeshift_f64_ranged"] = prev["sglm_timeshift_f64_ranged"]
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    if len(sys.argv) > 4:
        finish(sys.argv[2], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
