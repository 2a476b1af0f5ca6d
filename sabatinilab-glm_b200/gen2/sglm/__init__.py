"""Second-generation package layout of the reference (`sglm/sglm/`): put this directory's parent
(`sabatinilab-glm_b200/gen2`) on sys.path and `from sglm.models import sglm, sglm_cv, split_data, eval`,
`from sglm.features import sglm_pp, setup_model_fit`, `from sglm.data import save_results` work as they do with the
reference's cookiecutter package — backed by the same B200 kernels as the flat modules.  (The flat module `sglm.py`
of the first generation and this package share the name `sglm`, exactly as in the reference, where they live in
different trees: `backend/` vs `sglm/`.)"""
import os
import sys

_FLAT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _FLAT not in sys.path:
    sys.path.append(_FLAT)       # after this package, so that `sglm` keeps meaning the package
