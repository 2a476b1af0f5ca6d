"""sglm.features.setup_model_fit — reference sglm/sglm/features/setup_model_fit.py:43-96 (`timeshift_vals_by_dict`,
`X_cols_dict_to_default`)."""
from setup_model_fit import *  # noqa: F401,F403
from setup_model_fit import X_cols_dict_to_default, timeshift_vals_by_dict  # noqa: F401
