"""sglm.models.sglm — reference sglm/sglm/models/sglm.py: class GLM (:25-436, incl. the `closed_form` least-squares
branch :263-293), calc_R2 (:438-459) and fit_GLM (:461-483, moved here from the first generation's sglm_ez)."""
from _glm import *  # noqa: F401,F403
from _glm import GLM, NotYetImplementedError, calc_R2  # noqa: F401


def fit_GLM(X, y, model_name='Gaussian', *args, **kwargs):
    """Fit one GLM on DataFrame X / Series y (sglm/sglm/models/sglm.py:461-483)."""
    import sglm_ez
    return sglm_ez.fit_GLM(X, y, model_name, *args, **kwargs)
