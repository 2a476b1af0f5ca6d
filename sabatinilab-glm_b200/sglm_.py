"""sglm_ — the reference ships `backend/sglm_.py` as a byte-identical copy of
`backend/sglm.py` (sglm_cv.py:3 and sglm_ez.py:4 import this name, the tests import
`sglm`); both names resolve to the same objects here."""
from _glm import *  # noqa: F401,F403
from _glm import GLM, NotYetImplementedError, calc_R2  # noqa: F401
