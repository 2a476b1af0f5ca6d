"""sglm — drop-in for the reference module `backend/sglm.py` (class GLM, calc_R2).  The implementation lives in
`_glm.py` so that the second-generation package layout (`gen2/sglm/models/sglm.py`, reference sglm/sglm/models/sglm.py)
can share it without a name clash between this flat module and that package."""
from _glm import *  # noqa: F401,F403
from _glm import GLM, NotYetImplementedError, calc_R2  # noqa: F401
