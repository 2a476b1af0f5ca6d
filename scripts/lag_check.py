"""GPU check of the lag-design statistics: G from the base signals == G from the built design."""
import os, sys, time
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_pp, sglm_cv
T, P = int(os.environ.get("T", 60000)), 12
shifts = [0] + [s for s in range(-7, 9) if s != 0]
X0 = synth_data.synth_base(T, P, 5)
d = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts, device=True).dropna()
rec = d.lag_recipe()
print("rows", rec.lo, rec.hi, "window", rec.window()[:2])
Xd = rec.tensor()
n, C = Xd.shape
y = torch.randn(n, dtype=torch.float64, device="cuda")
Yd = y[:, None].contiguous()
folds = synth_data.synth_folds(n, 4, 3, group=500)
rows = [None] + [torch.from_numpy(b).cuda() for _, b in folds]
G0, s0 = eng.suffstats_tc(Xd, Yd, rows)
G1, s1 = eng.suffstats_tc(rec, Yd, rows)
print("planes", s0.sum(), s1.sum(), "equal colS", np.array_equal(s0, s1))
print("max abs diff", float((G0 - G1).abs().max()), "bit-equal", bool(torch.equal(G0, G1)))
eng.TC_CELLS = False
G2, _ = eng.suffstats_tc(rec, Yd, rows)
print("no-cells: bit-equal", bool(torch.equal(G0, G2)))
