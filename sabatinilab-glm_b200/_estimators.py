"""
Estimator objects behind `GLM.model` — the GPU counterparts of the scikit-learn classes
the reference instantiates (backend/sglm.py:2, :95-130): LinearRegression, Ridge, Lasso,
ElasticNet, TweedieRegressor(power=1).  Same constructor keywords (unknown keywords raise
TypeError exactly as scikit-learn does, which is how stale kwargs such as `reg_lambda`
fail in the reference), same fitted attributes (`coef_`, `intercept_`, `n_iter_`,
`dual_gap_`), same `fit / predict / score`.  All arithmetic runs in libsglm_b200.so.
"""
import warnings

import numpy as np

import _engine as eng


class ConvergenceWarning(UserWarning):
    """Mirror of sklearn.exceptions.ConvergenceWarning (a warning, not an error)."""


class _Base:
    kind = None
    _link = 0

    def get_params(self, deep=True):
        import inspect
        return {k: getattr(self, k) for k in inspect.signature(type(self).__init__).parameters if k != "self"}

    def set_params(self, **params):
        for k, v in params.items():
            if k not in self.get_params():
                raise ValueError(f"Invalid parameter {k!r} for estimator {type(self).__name__}")
            setattr(self, k, v)
        return self

    def __repr__(self):
        import inspect
        sig = inspect.signature(type(self).__init__).parameters
        diff = [f"{k}={getattr(self, k)!r}" for k, p in sig.items()
                if k != "self" and getattr(self, k) != p.default]
        return f"{type(self).__name__}({', '.join(diff)})"

    # -- shared fit path: statistics -> centred problem -> solver -> intercept
    def _check_supported(self):
        if getattr(self, "positive", False):
            raise NotImplementedError("positive=True is not part of the sGLM hot path")
        if getattr(self, "selection", "cyclic") != "cyclic":
            raise NotImplementedError("selection='random' is not part of the sGLM hot path")

    def _spec(self, problem):
        raise NotImplementedError

    def fit(self, X, y, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is never passed by the reference (backend/sglm.py:241)")
        self._check_supported()
        Xd = eng.device_matrix(X)
        yd = eng.device_vector(y)
        if Xd.shape[0] != yd.shape[0]:
            raise ValueError(f"Found input variables with inconsistent numbers of samples: "
                             f"[{Xd.shape[0]}, {yd.shape[0]}]")
        if Xd.shape[0] < 1:
            raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")
        C = Xd.shape[1]
        G = eng.suffstats(Xd, yd[:, None])
        _require_finite(G)
        prob = eng.center(G[0], None, C, 1, 0, self.fit_intercept)
        eng.fetch_scalars([prob])
        spec = self._spec(prob)
        W, info, status = eng.solve_models([spec], C)
        b, _ = eng.finalize(W, C, 1, [spec])
        self.coef_ = W[0, :C].cpu().numpy()
        self.intercept_ = float(b[0].item()) if self.fit_intercept else 0.0
        self.n_features_in_ = C
        self._after_fit(info[0], int(status[0]))
        if not (np.isfinite(self.coef_).all() and np.isfinite(self.intercept_)):
            if status[0] == 2:
                raise np.linalg.LinAlgError("Matrix is singular: X'X + alpha*I is not positive definite")
            raise ValueError("Coordinate descent iterations resulted in non-finite parameter values. The input "
                             "data may contain large values and need to be preprocessed.")
        return self

    def _after_fit(self, info, status):
        pass

    def predict(self, X):
        Xd = eng.device_matrix(X)
        out = eng.predict(Xd, self.coef_, self.intercept_, self._link)
        return out if eng.is_torch(eng._values(X)) else out.cpu().numpy()

    def score(self, X, y, sample_weight=None):
        Xd, yd = eng.device_matrix(X), eng.device_vector(y)
        s, _ = eng.score_sums(Xd, yd, self.coef_, self.intercept_, self._link)
        return eng.r2_from_sums(s)


def _require_finite(G):
    import torch
    if not bool(torch.isfinite(G).all().item()):
        raise ValueError("Input contains NaN, infinity or a value too large for dtype('float64').")


class LinearRegression(_Base):
    """OLS (reference: backend/sglm.py:96-101; sklearn _base.py:700-756)."""
    kind = "ols"

    def __init__(self, *, fit_intercept=True, copy_X=True, tol=1e-6, n_jobs=None, positive=False):
        self.fit_intercept, self.copy_X, self.tol, self.n_jobs, self.positive = \
            fit_intercept, copy_X, tol, n_jobs, positive

    def _spec(self, problem):
        return eng.ModelSpec(problem, "ols", tol=self.tol)      # tol = lstsq cond (_base.py:750-753)


class Ridge(_Base):
    """Ridge, alpha NOT scaled by n (reference: backend/sglm.py:102-105; sklearn _ridge.py:215-227)."""
    kind = "ridge"

    def __init__(self, alpha=1.0, *, fit_intercept=True, copy_X=True, max_iter=None, tol=1e-4,
                 solver="auto", positive=False, random_state=None):
        self.alpha, self.fit_intercept, self.copy_X, self.max_iter, self.tol = \
            alpha, fit_intercept, copy_X, max_iter, tol
        self.solver, self.positive, self.random_state = solver, positive, random_state

    def _spec(self, problem):
        if self.solver not in ("auto", "cholesky"):
            raise NotImplementedError(f"Ridge solver={self.solver!r}: only the reference's default (cholesky) is built")
        if self.alpha < 0:
            raise ValueError("The 'alpha' parameter of Ridge must be a float in the range [0.0, inf).")
        return eng.ModelSpec(problem, "ridge", alpha=self.alpha)


class ElasticNet(_Base):
    """ElasticNet by cyclic coordinate descent (reference: backend/sglm.py:109-110;
    sklearn _coordinate_descent.py:1095-1281, _cd_fast.pyx:243-506)."""
    kind = "enet"

    def __init__(self, alpha=1.0, *, l1_ratio=0.5, fit_intercept=True, precompute=False, max_iter=1000,
                 copy_X=True, tol=1e-4, warm_start=False, positive=False, random_state=None,
                 selection="cyclic"):
        self.alpha, self.l1_ratio, self.fit_intercept, self.precompute = alpha, l1_ratio, fit_intercept, precompute
        self.max_iter, self.copy_X, self.tol, self.warm_start = max_iter, copy_X, tol, warm_start
        self.positive, self.random_state, self.selection = positive, random_state, selection

    def _spec(self, problem):
        init = getattr(self, "coef_", None) if self.warm_start else None
        if self.alpha == 0:
            warnings.warn("With alpha=0, this algorithm does not converge well. You are advised to use the "
                          "LinearRegression estimator", stacklevel=3)
        return eng.ModelSpec(problem, "enet", alpha=self.alpha, l1_ratio=self.l1_ratio, max_iter=self.max_iter,
                             tol=self.tol, coef_init=init)

    def _after_fit(self, info, status):
        self.dual_gap_ = float(info[0])
        self.n_iter_ = int(info[2])
        if status == 1:
            warnings.warn(f"Objective did not converge. You might want to increase the number of iterations, "
                          f"check the scale of the features or consider increasing regularisation. "
                          f"Duality gap: {info[0]:.6e}, tolerance: {info[1]:.3e}", ConvergenceWarning, stacklevel=3)


class Lasso(ElasticNet):
    """Lasso = ElasticNet(l1_ratio=1) (reference: backend/sglm.py:106-108)."""
    kind = "lasso"

    def __init__(self, alpha=1.0, *, fit_intercept=True, precompute=False, copy_X=True, max_iter=1000,
                 tol=1e-4, warm_start=False, positive=False, random_state=None, selection="cyclic"):
        super().__init__(alpha=alpha, l1_ratio=1.0, fit_intercept=fit_intercept, precompute=precompute,
                         copy_X=copy_X, max_iter=max_iter, tol=tol, warm_start=warm_start, positive=positive,
                         random_state=random_state, selection=selection)

    def get_params(self, deep=True):
        p = super().get_params(deep)
        p.pop("l1_ratio", None)
        return p


class TweedieRegressor(_Base):
    """Poisson GLM with log link (reference: backend/sglm.py:112-115 builds
    TweedieRegressor(power=1); sklearn _glm/glm.py:185-339).  Fitted by IRLS/Newton on the
    GPU to the optimum of  mean(mu - y*eta) + alpha/2 |w|^2 ; `score` is D^2."""
    kind = "poisson"
    _link = 1

    def __init__(self, *, power=0.0, alpha=1.0, fit_intercept=True, link="auto", solver="lbfgs",
                 max_iter=100, tol=1e-4, warm_start=False, verbose=0):
        self.power, self.alpha, self.fit_intercept, self.link = power, alpha, fit_intercept, link
        self.solver, self.max_iter, self.tol, self.warm_start, self.verbose = solver, max_iter, tol, warm_start, verbose

    def _check_family(self):
        if self.power != 1 or self.link not in ("auto", "log"):
            raise NotImplementedError("only the Poisson family with log link (power=1) is built on the GPU path; "
                                      "Gamma/Tweedie(power!=1) are outside the hot-path scope")

    def fit(self, X, y, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is never passed by the reference")
        self._check_family()
        Xd, yd = eng.device_matrix(X), eng.device_vector(y)
        import torch
        if bool((yd < 0).any().item()):
            raise ValueError("Some value(s) of y are out of the valid range of the loss 'HalfPoissonLoss'.")
        if not bool(torch.isfinite(Xd).all().item()):
            raise ValueError("Input X contains NaN or infinity.")
        ci = getattr(self, "coef_", None) if self.warm_start else None
        ii = getattr(self, "intercept_", None) if self.warm_start else None
        self.coef_, self.intercept_, self.n_iter_ = eng.poisson_irls(
            Xd, yd, self.alpha, self.fit_intercept, None, self.max_iter, self.tol, ci, ii)
        self.n_features_in_ = Xd.shape[1]
        return self

    def score(self, X, y, sample_weight=None):
        self._check_family()
        Xd, yd = eng.device_matrix(X), eng.device_vector(y)
        s, _ = eng.score_sums(Xd, yd, self.coef_, self.intercept_, 1)
        return eng.poisson_d2_from_sums(s)
