"""sglm.models.train_model — reference sglm/sglm/models/train_model.py:3-44 (frame adaptors around the fit)."""


def get_x(df, x_cols, keep_rows=None):
    if keep_rows is not None:
        df = df[keep_rows]
    return df[x_cols]


def get_y(df, y_col, keep_rows=None):
    if keep_rows is not None:
        df = df[keep_rows]
    return df[y_col]


def get_xy_all_noniti(df, prediction_X_cols_sftd, y_col, noniticol='wi_trial_keep'):
    return (get_x(df, prediction_X_cols_sftd), get_y(df, y_col),
            get_x(df, prediction_X_cols_sftd, keep_rows=df[noniticol]), get_y(df, y_col, keep_rows=df[noniticol]))


def setup_glmsave(glmsave, prefix, filename, neg_order, pos_order, X_cols_all, folds, pholdout, pgss, gssid=None):
    glmsave.set_uid(prefix)
    glmsave.set_filename(filename)
    glmsave.set_timeshifts(neg_order, pos_order)
    glmsave.set_X_cols(X_cols_all)
    glmsave.set_gss_info(folds, pholdout, pgss, gssid=None)        # the reference drops gssid here (:43)
