"""Timeline of BASELINE config 1 (one ElasticNet fit, 100k x 410, through GLM.fit)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm, sglm_pp, _sglm_native as nat
T, P = 100_000, 10
shifts = [0] + [s for s in range(-20, 21) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 1)).cuda()
beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 1)).cuda()
def design():
    return sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[20:T - 20]
X = design()
torch.manual_seed(1)
s = X @ beta; y = s + 1.5 * s.std() * torch.randn_like(s); y = ((y - y.mean()) / y.std()).contiguous()
def step():
    return sglm.GLM("Gaussian", alpha=0.01, l1_ratio=0.5).fit(design(), y)
step(); step(); torch.cuda.synchronize()
nat.enable_timing(True); nat.collect_timing()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
nat.collect_timing()
print(f"step {e0.elapsed_time(e1):.2f} ms")
prev = 0.0
for name, a, b in sorted(nat.last_intervals, key=lambda t: t[1]):
    print(f"{a:8.3f} -> {b:8.3f}  ({b - a:7.3f} ms, gap before {a - prev:6.3f})  {name}")
    prev = max(prev, b)
