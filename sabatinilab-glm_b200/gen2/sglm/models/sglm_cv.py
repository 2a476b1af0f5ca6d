"""sglm.models.sglm_cv — reference sglm/sglm/models/sglm_cv.py: simple_cv_fit (:18-62), cv_idx_by_timeframe
(:64-88), SGLM_worker (:90-129), cv_glm_single_params (:131-255), cv_glm_mult_params (:257-358),
generate_mult_params (:360-383).  Same arithmetic as the first generation (the "L2 1/(2N) ratio fix" :222-224 is
commented out in the reference); its changes — 4 parameter-set workers (:322), queue time-outs (:103-108), no PCA
prefit (:299-301) — are scheduling details that the batched B200 plan replaces (every parameter set and fold of a
call runs in the same launches), so both generations resolve to one implementation."""
from sglm_cv import (SGLM_worker, cv_glm_mult_params, cv_glm_mult_params_sessions, cv_glm_single_params,  # noqa: F401
                     generate_mult_params)
from sglm_ez import cv_idx_by_timeframe, simple_cv_fit  # noqa: F401
