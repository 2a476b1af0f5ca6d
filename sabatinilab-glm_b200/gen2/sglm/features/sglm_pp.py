"""sglm.features.sglm_pp — reference sglm/sglm/features/sglm_pp.py: the first generation's sglm_pp plus the
sglm_ez timeshift helpers merged in (timeshift_cols, add_timeshifts_to_col_list, ...)."""
from sglm_pp import *  # noqa: F401,F403
from sglm_pp import (DeviceDesign, bucket_ids_by_timeframe, cv_idx_from_bucket_ids, detrend_data, diff,  # noqa: F401
                     get_column_nums, timeshift, timeshift_multiple, zscore)
from sglm_ez import (add_timeshifts_by_sl_to_col_list, add_timeshifts_to_col_list, diff_cols, timeshift_cols,  # noqa: F401
                     timeshift_cols_by_signal_length)
