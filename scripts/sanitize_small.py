"""Small run through every non-tcgen05 kernel (for `compute-sanitizer --tool memcheck`)."""
import os, sys
os.environ["SGLM_GRAM"] = "dmma"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import numpy as np, torch
import synth_data, sglm, sglm_pp, sglm_cv
X0 = synth_data.synth_base(3000, 5, 1)
shifts = [0, -3, -2, -1, 1, 2, 3]
Xd = sglm_pp.timeshift_multiple(X0, shift_amt_list=shifts)[3:-3]
sglm_pp.timeshift(np.random.rand(50, 3000), shift_inx=[0, 2999], shift_amt=7)          # direct kernel
sglm_pp.timeshift(np.random.rand(9, 3), shift_amt=20)                                  # oversize shift
y = synth_data.synth_response(Xd, synth_data.synth_kernels(5, shifts, 1), 1)
cv = synth_data.synth_folds(Xd.shape[0], 3, 1, group=100)
grid = [dict(alpha=a, l1_ratio=l, max_iter=200) for a in (0, 1e-3, 1e-1) for l in (0, 0.5, 1)]
r = sglm_cv.cv_glm_mult_params(Xd, y, cv, "Gaussian", grid, score_method="r2")
g = r["best_model"]; g.predict(Xd); g.r2_score(Xd, y); g.neg_mse_score(Xd, y); g.get_residuals(Xd, y)
Xw = np.random.default_rng(0).standard_normal((700, 333)); yw = Xw[:, 0] + np.random.default_rng(1).standard_normal(700)
sglm.GLM("Gaussian", alpha=0.05, l1_ratio=0.5).fit(Xw, yw)                             # odd C, 2 panel warps
sglm.GLM("Gaussian", alpha=1.0, l1_ratio=0, max_iter=5).fit(Xw, yw)                    # cholesky, ragged tiles
yp = synth_data.synth_response(Xd, synth_data.synth_kernels(5, shifts, 1), 1, poisson=True)
sglm.GLM("Poisson", alpha=0.01).fit(Xd, yp)
torch.cuda.synchronize(); print("sanitize_small done")
