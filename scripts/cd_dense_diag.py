"""In-kernel cost of the dense register chain: clocks inside cd_dense_chain / dense blocks, and of the whole register phase / blocks."""
import os, sys
sys.argv = [sys.argv[0]]
exec(open(os.path.join(os.path.dirname(__file__), "cd_timers.py")).read().split("sub = int(")[0])
models = ms[::8]
for plan in ["4x4@0.3,0x0", "2x4@0.3,0x0", "4x2@0.3,0x0"]:
    out = {}
    for sel in (8, 9, 10):
        eng.CD_PLAN, eng.CD_DEBUG_TIMER = plan, sel
        Wd, info, st = eng.solve_models(models, C)
        h = int(np.argmax(info[:, 4]))
        out[sel] = (info[h, 5], info[h, 4])
    print(f"plan {plan}: heaviest model blocks {out[9][1]:.0f}, dense blocks {out[9][0]:.0f}, clocks in the dense chain per dense block {out[8][0] / max(out[9][0], 1):.0f}, "
          f"register phase clocks per block {out[10][0] / out[10][1]:.0f}", flush=True)
