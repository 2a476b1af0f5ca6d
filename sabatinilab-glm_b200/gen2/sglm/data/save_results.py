"""sglm.data.save_results — reference sglm/sglm/data/save_results.py:11-73 == sglm_save.py:7-68 (`GLM_data`)."""
from sglm_save import GLM_data, save_model_arrays  # noqa: F401
