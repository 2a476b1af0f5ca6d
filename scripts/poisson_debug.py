import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sabatinilab-glm_b200"))
import torch
import synth_data, _engine as eng, sglm_cv
from oracle import sglm_oracle as orc
T, P = 12_000, 20
shifts = [0] + [s for s in range(-20, 20) if s != 0]
X0 = synth_data.synth_base(T, P, 404)
Xd = orc.timeshift_multiple(X0, shift_amt_list=shifts)
Xd = Xd[~np.isnan(Xd).any(axis=1)]
y = synth_data.synth_response(Xd, synth_data.synth_kernels(P, shifts, 404), 404, poisson=True)
cv_idx = synth_data.synth_folds(Xd.shape[0], 3, 404, group=500)
grid = [dict(alpha=a, model_name="Poisson") for a in (1e-3, 1e-2, 0.1, 1.0)] + [
    dict(alpha=0.05, fit_intercept=False, model_name="Poisson"), dict(alpha=0.02, roll=5, model_name="Poisson")]
want = orc.cv_glm_mult_params(Xd, y, cv_idx, "Poisson", [dict(k) for k in grid], score_method="r2")
for use_tc in (False, True):
    eng.POISSON_TC = use_tc
    got = sglm_cv.cv_glm_mult_params(Xd, y, cv_idx, "Poisson", [dict(k) for k in grid], score_method="r2")
    print("use_tc", use_tc, eng.last_poisson_batch)
    for a, b in zip(got["full_cv_results"], want["full_cv_results"]):
        ce = [float(np.max(np.abs(a["cv_coefs"][:, k] - b["cv_coefs"][:, k])) / np.max(np.abs(b["cv_coefs"][:, k]))) for k in range(3)]
        print(a["glm_kwargs"], "status", a["_fit_info"]["status"], "n_iter", a["_fit_info"]["n_iter"], "coef err", ["%.1e" % e for e in ce],
              "refit err %.1e" % (np.max(np.abs(a["model"].coef_ - b["model"].coef_)) / np.max(np.abs(b["model"].coef_))),
              "icpt", a["cv_intercepts"], b["cv_intercepts"])
        print("   R2", a["cv_R2_score"], b["cv_R2_score"], "test", a["cv_scores_test"], b["cv_scores_test"], "train", a["cv_scores_train"], b["cv_scores_train"],
              "mse", a["cv_mse_score"], b["cv_mse_score"])
