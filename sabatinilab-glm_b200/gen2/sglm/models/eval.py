"""sglm.models.eval — reference sglm/sglm/models/eval.py: calc_l1 / calc_l2 (:5-9), print_best_model_info (:12-45),
training_fit_holdout_score (:48-72); plus `holdout_scores` (extension): the per-model hold-out scoring loop of the
drivers (er_refactored_from_scratch_cleanup.py:528-550) as one batched pass."""
import time

import numpy as np

import sglm_ez
from sglm_ez import training_fit_holdout_score  # noqa: F401
from sglm_ez import holdout_scores  # noqa: F401


def calc_l1(coeffs):
    return np.sum(np.abs(coeffs))


def calc_l2(coeffs):
    return np.sum(np.square(coeffs))


def print_best_model_info(X_setup, best_score, best_params, best_model, start, show_non_zero_coefs=False):
    print()
    print('---')
    print()
    if show_non_zero_coefs:
        print('Non-Zero Coeffs:')
        for ic, coef in enumerate(best_model.coef_):
            if np.abs(coef) > 1e-10:
                print(f'> {coef}: {X_setup.columns[ic]}')
    print(f'Best Score: {best_score}')
    print(f'Best Params: {best_params}')
    print(f'Best Model: {best_model}')
    print(f'Best Model — Intercept: {best_model.intercept_}')
    print(f'Overall RunTime: {time.time() - start}')
    print()
