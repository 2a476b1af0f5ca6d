import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/sabatinilab-glm_b200'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import _engine as eng
rng = np.random.default_rng(0)
T, C = 20000, 300
X = rng.standard_normal((T, C))
X[:, :100] = (rng.random((T, 100)) < 0.02).astype(float)         # 0/1 events
X[:, 100:120] = rng.integers(-5, 6, (T, 20)).astype(float)       # small integers
X[:, 120:140] = rng.integers(0, 2**20, (T, 20)) / 1024.0         # fixed point
X[:, 140] = 0.0
X[:, 141] = 1e-310                                                 # subnormal
X[:, 142] = np.float32(rng.standard_normal(T)).astype(np.float64) # float32-representable
X[:, 143] = 3.0
X[:, 144] = rng.standard_normal(T) * 1e12
y = rng.standard_normal((T, 1))
Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
_, s1 = eng.suffstats_tc(Xd, Yd, [None])
os.environ["SGLM_TUNING"] = "1"
os.environ["SGLM_TC_DIGIT_PASS"] = "1"
_, s2 = eng.suffstats_tc(Xd, Yd, [None])
print("one-pass == digit pass:", np.array_equal(s1, s2), "planes", s1.sum(), s2.sum())
print("diff cols:", np.flatnonzero(s1 != s2)[:20], s1[s1 != s2][:20], s2[s1 != s2][:20])
print(s1[[0, 100, 120, 140, 141, 142, 143, 144, 200]])
