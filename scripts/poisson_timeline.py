import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sabatinilab-glm_b200')
import numpy as np, torch
import synth_data, sglm_pp, _engine as eng, _sglm_native as nat
T, P = 1_000_000, 20
shifts = [0] + [s for s in range(-20, 20) if s != 0]
X0 = torch.from_numpy(synth_data.synth_base(T, P, 5)).cuda()
beta = torch.from_numpy(synth_data.synth_kernels(P, shifts, 5)).cuda()
X = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[19:T - 20]
s = X @ beta; z = (s - s.mean()) / s.std()
torch.manual_seed(0); y = torch.poisson(torch.exp(0.3 * z - 1.0)).contiguous()
eng.poisson_irls(X, y, 0.01)
torch.cuda.synchronize()
nat.enable_timing(True); nat.collect_timing()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); w, b, n = eng.poisson_irls(X, y, 0.01); e1.record(); torch.cuda.synchronize()
nat.collect_timing()
print("fit", e0.elapsed_time(e1), "ms", n, "iterations")
prev = 0
for name, a, bb in sorted(nat.last_intervals, key=lambda t: t[1]):
    print(f"{a:8.2f} -> {bb:8.2f} ({bb-a:6.2f} ms, gap {a-prev:6.2f}) {name}")
    prev = max(prev, bb)
