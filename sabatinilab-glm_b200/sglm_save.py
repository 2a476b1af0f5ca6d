"""
sglm_save — the on-disk result formats of the reference, write-compatible so that the lab's
downstream plotting notebooks keep working:

  * `GLM_data` (sglm_save.py:7-68 == sglm/sglm/data/save_results.py:11-73): a pickled result container.  Field
    names, defaults and the (odd) load/save asymmetry are kept: `save` pickles the whole object, `load` assigns the
    unpickled object to `.data` (sglm_save.py:13-31).  The fitted models inside are this package's `GLM` objects:
    plain Python objects holding numpy arrays (no device state), so they unpickle on a machine without a GPU; their
    `predict / score` need the B200 library, their `coef_ / intercept_` do not.
  * `save_model_arrays`: the per-model `np.save` files the drivers write
    (er_refactored_from_scratch_cleanup.py:528-537):
        {folder}/coeffs/{run_id}_{k1}_{v1}_{k2}_{v2}..._{coef_basename}.npy
        {folder}/intercepts/{run_id}_..._{intercept_basename}.npy
"""
import os
import pickle
from os.path import exists

import numpy as np


class GLM_data():
    def __init__(self, file_dir, filename):
        self.file_dir = file_dir
        self.filename = filename
        self.data = {}
        self.data['fit_results'] = []

    def save(self, overwrite=False):
        path_to_file = self.file_dir + '/' + self.filename
        if not exists(path_to_file) or overwrite:
            with open(path_to_file, 'wb') as file_save:
                pickle.dump(self, file_save)
            print('SGLM file saved to: ' + path_to_file)
        else:
            print('File already exists. Set overwrite=True to overwrite.')

    def load(self):
        path_to_file = self.file_dir + '/' + self.filename
        if not exists(path_to_file):
            print('File does not exist.')
            return
        with open(path_to_file, 'rb') as file_load:
            self.data = pickle.load(file_load)

    def set_uid(self, uid):
        self.data['uid'] = uid

    def set_filename(self, filename):
        self.data['filename'] = filename

    def set_basedata(self, basedata):
        self.data['basedata'] = basedata

    def set_X_cols(self, X_cols):
        self.data['X_cols'] = X_cols

    def set_gss_info(self, folds, pholdout, pgss, gssid=None):
        self.data['gss_info'] = {'folds': folds, 'pholdout': pholdout, 'pgss': pgss, 'gssid': gssid}

    def set_timeshifts(self, negorder, posorder):
        self.data['negorder'] = negorder
        self.data['posorder'] = posorder

    def append_fit_results(self, response_col, hyperparams, glm_model=None, scores=None, dropped_cols=[], gssids=None):
        for score_id in ['tr_witi', 'tr_noiti', 'gss_witi', 'gss_noiti', 'holdout_witi', 'holdout_noiti']:
            if score_id not in scores:
                scores[score_id] = None
        self.data['fit_results'].append({'response_col': response_col, 'hyperparams': hyperparams,
                                         'glm_model_gss': glm_model, 'dropped_cols': dropped_cols, 'scores': scores,
                                         'gss_mse': None, 'refit_mse': None, 'gssids': gssids})


def model_file_stem(run_id, glm_kwargs):
    """`{run_id}_{k}_{v}_...` in the dict's key order (er_refactored_from_scratch_cleanup.py:530-535)."""
    kwarg_info = "_".join([f"{k}_{glm_kwargs[k]}" for k in glm_kwargs])
    return f'{run_id}_{kwarg_info}'


def save_model_arrays(cv_results, run_id, all_models_folder, model_c_basename='coeffs', model_i_basename='intercept'):
    """One coefficient and one intercept `.npy` per parameter set of `cv_results['full_cv_results']`, named as the
    drivers name them.  Returns the list of (coef path, intercept path)."""
    os.makedirs(f'{all_models_folder}/coeffs', exist_ok=True)
    os.makedirs(f'{all_models_folder}/intercepts', exist_ok=True)
    out = []
    for fitted in cv_results['full_cv_results']:
        model = fitted['model']
        std_name = model_file_stem(run_id, fitted['glm_kwargs'])
        pc = f'{all_models_folder}/coeffs/{std_name}_{model_c_basename}.npy'
        pi = f'{all_models_folder}/intercepts/{std_name}_{model_i_basename}.npy'
        np.save(pc, model.coef_)
        np.save(pi, model.intercept_)
        out.append((pc, pi))
    return out
