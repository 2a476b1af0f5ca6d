from . import sglm, sglm_cv, split_data, eval, train_model  # noqa: F401
