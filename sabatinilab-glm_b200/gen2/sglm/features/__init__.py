from . import sglm_pp, setup_model_fit  # noqa: F401
