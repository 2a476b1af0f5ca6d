/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (sabatinilab-glm_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as
 * the checker or the reported CPU baseline.
 *
 * Plain-C restatement of the coordinate-descent solvers that the reference
 * reaches through  backend/sglm.py:106-110  (Lasso / ElasticNet construction)
 * and  backend/sglm.py:241  (self.model.fit).  The arithmetic itself lives in
 * the un-vendored third-party dependency scikit-learn (requirements.txt:7 pins
 * scikit_learn==0.24.2; the container has 1.9.0).  The algorithm restated here
 * is the one published in
 *     sklearn/linear_model/_cd_fast.pyx:243-506   enet_coordinate_descent
 *     sklearn/linear_model/_cd_fast.pyx:1006-1290 gap_enet_gram +
 *                                                 enet_coordinate_descent_gram
 * (1.9.0 numbering), i.e. cyclic coordinate descent with soft-thresholding,
 * the  d_w_max / w_max <= tol  trigger, the duality-gap stop  gap <= tol*||y||^2
 * and (1.9 only) gap-safe screening.  Pinned by tests/test_oracle_pins.py
 * against scikit-learn run in this container (fixtures under tests/golden/).
 *
 * Build:  make -C oracle     ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double ddot(int n, const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static void daxpy(int n, double a, const double *x, double *y) {
    for (int i = 0; i < n; ++i) y[i] += a * x[i];
}

static double fsign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

/* _cd_fast.pyx:133-156  dual_gap_formulation_A */
static double gap_formulation_A(double alpha, double beta, double w_l1, double w_l22,
                                double R2, double Ry, double dual_norm) {
    double primal = 0.5 * (R2 + beta * w_l22) + alpha * w_l1;
    double scale = dual_norm > alpha ? alpha / dual_norm : 1.0;
    double dual = -0.5 * scale * scale * (R2 + beta * w_l22) + scale * Ry;
    return primal - dual;
}

/* _cd_fast.pyx:159-240  gap_enet  (dense X, column-major n x p) */
static double gap_enet(int n, int p, const double *w, double alpha, double beta,
                       const double *X, const double *y, const double *R, double *XtA,
                       double *dual_norm_out) {
    double w_l22 = 0.0, R2, Ry = 0.0, dual_norm, w_l1 = 0.0;
    if (beta > 0) w_l22 = ddot(p, w, w);
    R2 = ddot(n, R, R);
    if (!(alpha == 0 && beta == 0)) Ry = ddot(n, R, y);
    if (alpha == 0) {
        for (int j = 0; j < p; ++j) XtA[j] = ddot(n, X + (size_t)j * n, R);
        dual_norm = ddot(p, XtA, XtA);
        *dual_norm_out = dual_norm;
        if (beta == 0) return dual_norm;
        return R2 + 0.5 * beta * w_l22 - Ry + 1.0 / (2.0 * beta) * dual_norm;
    }
    dual_norm = 0.0;
    for (int j = 0; j < p; ++j) {
        XtA[j] = ddot(n, X + (size_t)j * n, R) - beta * w[j];
        double a = fabs(XtA[j]);
        if (a > dual_norm) dual_norm = a;
        w_l1 += fabs(w[j]);
    }
    *dual_norm_out = dual_norm;
    return gap_formulation_A(alpha, beta, w_l1, w_l22, R2, Ry, dual_norm);
}

/*
 * _cd_fast.pyx:243-506.  X column-major (n x p), already centred by the caller
 * when fit_intercept (sklearn/linear_model/_base.py:832-928 _pre_fit).
 * w is in/out (warm start).  Returns 0, or 1 when max_iter was exhausted
 * without the gap test passing (sklearn: ConvergenceWarning).
 */
int sglm_oracle_enet_cd(double *w, double alpha, double beta, const double *X,
                        const double *y, int n, int p, int max_iter, double tol,
                        int do_screening, double *gap_out, double *tol_out,
                        int *n_iter_out) {
    double *norm2 = (double *)malloc(sizeof(double) * p);
    double *R = (double *)malloc(sizeof(double) * n);
    double *XtA = (double *)malloc(sizeof(double) * p);
    int *active = (int *)malloc(sizeof(int) * p);
    unsigned char *excluded = (unsigned char *)calloc(p, 1);
    double gap = tol + 1.0, d_w_tol = tol, dual_norm = 0.0;
    int n_active = p, n_iter = 0, converged = 0;

    for (int j = 0; j < p; ++j) norm2[j] = ddot(n, X + (size_t)j * n, X + (size_t)j * n);
    if (alpha == 0) do_screening = 0;

    memcpy(R, y, sizeof(double) * n);
    for (int j = 0; j < p; ++j)
        if (w[j] != 0) daxpy(n, -w[j], X + (size_t)j * n, R);
    tol *= ddot(n, y, y);

    gap = gap_enet(n, p, w, alpha, beta, X, y, R, XtA, &dual_norm);
    if (gap <= tol) { n_iter = 0; converged = 1; goto done; }

    if (do_screening) {
        n_active = 0;
        for (int j = 0; j < p; ++j) {
            if (norm2[j] == 0) { w[j] = 0; excluded[j] = 1; continue; }
            double Xj_theta = XtA[j] / fmax(alpha, dual_norm);
            double d_j = (1 - fabs(Xj_theta)) / sqrt(norm2[j] + beta);
            if (d_j <= sqrt(2 * gap) / alpha) {
                active[n_active++] = j; excluded[j] = 0;
            } else {
                if (w[j] != 0) { daxpy(n, w[j], X + (size_t)j * n, R); w[j] = 0; }
                excluded[j] = 1;
            }
        }
    }

    for (n_iter = 0; n_iter < max_iter; ++n_iter) {
        double w_max = 0.0, d_w_max = 0.0;
        for (int f = 0; f < n_active; ++f) {
            int j = do_screening ? active[f] : f;
            if (norm2[j] == 0.0) continue;
            double w_j = w[j];
            const double *Xj = X + (size_t)j * n;
            double tmp = ddot(n, Xj, R) + w_j * norm2[j];
            w[j] = fsign(tmp) * fmax(fabs(tmp) - alpha, 0) / (norm2[j] + beta);
            if (w[j] != w_j) daxpy(n, w_j - w[j], Xj, R);
            double d_w_j = fabs(w[j] - w_j);
            d_w_max = fmax(d_w_max, d_w_j);
            w_max = fmax(w_max, fabs(w[j]));
        }
        if (w_max == 0.0 || d_w_max / w_max <= d_w_tol || n_iter == max_iter - 1) {
            gap = gap_enet(n, p, w, alpha, beta, X, y, R, XtA, &dual_norm);
            if (gap <= tol) { converged = 1; n_iter += 1; goto done; }
            if (do_screening) {
                n_active = 0;
                for (int j = 0; j < p; ++j) {
                    if (excluded[j]) continue;
                    double Xj_theta = XtA[j] / fmax(alpha, dual_norm);
                    double d_j = (1 - fabs(Xj_theta)) / sqrt(norm2[j] + beta);
                    if (d_j <= sqrt(2 * gap) / alpha) {
                        active[n_active++] = j; excluded[j] = 0;
                    } else {
                        if (w[j] != 0) { daxpy(n, w[j], X + (size_t)j * n, R); w[j] = 0; }
                        excluded[j] = 1;
                    }
                }
            }
        }
    }
    /* for/else: loop ended without break; sklearn returns n_iter + 1 == max_iter */
done:
    *gap_out = gap; *tol_out = tol; *n_iter_out = n_iter;
    free(norm2); free(R); free(XtA); free(active); free(excluded);
    return converged ? 0 : 1;
}

/* _cd_fast.pyx:1006-1092  gap_enet_gram */
static double gap_enet_gram(int p, const double *w, double alpha, double beta,
                            const double *Qw, const double *q, double y_norm2,
                            double *XtA, double *dual_norm_out) {
    double w_l22 = 0.0, q_dot_w, wQw, R2, Ry = 0.0, dual_norm, w_l1 = 0.0;
    if (beta > 0) w_l22 = ddot(p, w, w);
    q_dot_w = ddot(p, w, q);
    wQw = ddot(p, w, Qw);
    R2 = y_norm2 + wQw - 2.0 * q_dot_w;
    if (!(alpha == 0 && beta == 0)) Ry = y_norm2 - q_dot_w;
    if (alpha == 0) {
        for (int j = 0; j < p; ++j) XtA[j] = q[j] - Qw[j];
        dual_norm = ddot(p, XtA, XtA);
        *dual_norm_out = dual_norm;
        if (beta == 0) return dual_norm;
        return R2 + 0.5 * beta * w_l22 - Ry + 1.0 / (2.0 * beta) * dual_norm;
    }
    dual_norm = 0.0;
    for (int j = 0; j < p; ++j) {
        XtA[j] = q[j] - Qw[j] - beta * w[j];
        double a = fabs(XtA[j]);
        if (a > dual_norm) dual_norm = a;
        w_l1 += fabs(w[j]);
    }
    *dual_norm_out = dual_norm;
    return gap_formulation_A(alpha, beta, w_l1, w_l22, R2, Ry, dual_norm);
}

/*
 * _cd_fast.pyx:1095-1290.  Q is the (centred) Gram X'X, row-major p x p; q = X'y.
 * Used to pin the algorithm the CUDA kernel runs (same statistics, same sweep).
 * n_updates_out counts the coordinate updates that changed w (rows of Q read).
 */
int sglm_oracle_enet_cd_gram(double *w, double alpha, double beta, const double *Q,
                             const double *q, double y_norm2, int p, int max_iter,
                             double tol, int do_screening, double *gap_out,
                             double *tol_out, int *n_iter_out, long long *n_updates_out) {
    double *Qw = (double *)calloc(p, sizeof(double));
    double *XtA = (double *)calloc(p, sizeof(double));
    int *active = (int *)malloc(sizeof(int) * p);
    unsigned char *excluded = (unsigned char *)calloc(p, 1);
    double gap = tol + 1.0, d_w_tol = tol, dual_norm = 0.0, radius;
    int n_active = p, n_iter = 0, converged = 0;
    long long n_updates = 0;

    for (int j = 0; j < p; ++j)
        if (w[j] != 0) daxpy(p, w[j], Q + (size_t)j * p, Qw);
    if (alpha == 0) do_screening = 0;
    tol *= y_norm2;

    gap = gap_enet_gram(p, w, alpha, beta, Qw, q, y_norm2, XtA, &dual_norm);
    if (0 <= gap && gap <= tol) { converged = 1; n_iter = 0; goto done; }

    if (do_screening) {
        radius = sqrt(2 * fabs(gap)) / alpha;
        n_active = 0;
        for (int j = 0; j < p; ++j) {
            double Qjj = Q[(size_t)j * p + j];
            if (Qjj == 0) { w[j] = 0; excluded[j] = 1; continue; }
            double Xj_theta = XtA[j] / fmax(alpha, dual_norm);
            double d_j = (1 - fabs(Xj_theta)) / sqrt(Qjj + beta);
            if (d_j <= radius) {
                active[n_active++] = j; excluded[j] = 0;
            } else {
                if (w[j] != 0) { daxpy(p, -w[j], Q + (size_t)j * p, Qw); w[j] = 0; ++n_updates; }
                excluded[j] = 1;
            }
        }
    }

    for (n_iter = 0; n_iter < max_iter; ++n_iter) {
        double w_max = 0.0, d_w_max = 0.0;
        for (int f = 0; f < n_active; ++f) {
            int j = do_screening ? active[f] : f;
            double Qjj = Q[(size_t)j * p + j];
            if (Qjj == 0.0) continue;
            double w_j = w[j];
            double tmp = q[j] - Qw[j] + w_j * Qjj;
            w[j] = fsign(tmp) * fmax(fabs(tmp) - alpha, 0) / (Qjj + beta);
            if (w[j] != w_j) { daxpy(p, w[j] - w_j, Q + (size_t)j * p, Qw); ++n_updates; }
            double d_w_j = fabs(w[j] - w_j);
            if (d_w_j > d_w_max) d_w_max = d_w_j;
            if (fabs(w[j]) > w_max) w_max = fabs(w[j]);
        }
        if (w_max == 0.0 || d_w_max / w_max <= d_w_tol || n_iter == max_iter - 1) {
            gap = gap_enet_gram(p, w, alpha, beta, Qw, q, y_norm2, XtA, &dual_norm);
            if (gap <= tol) { converged = 1; n_iter += 1; goto done; }
            if (do_screening) {
                radius = sqrt(2 * fabs(gap)) / alpha;
                n_active = 0;
                for (int j = 0; j < p; ++j) {
                    if (excluded[j]) continue;
                    double Qjj = Q[(size_t)j * p + j];
                    double Xj_theta = XtA[j] / fmax(alpha, dual_norm);
                    double d_j = (1 - fabs(Xj_theta)) / sqrt(Qjj + beta);
                    if (d_j <= radius) {
                        active[n_active++] = j; excluded[j] = 0;
                    } else {
                        if (w[j] != 0) { daxpy(p, -w[j], Q + (size_t)j * p, Qw); w[j] = 0; ++n_updates; }
                        excluded[j] = 1;
                    }
                }
            }
        }
    }
done:
    *gap_out = gap; *tol_out = tol; *n_iter_out = n_iter;
    if (n_updates_out) *n_updates_out = n_updates;
    free(Qw); free(XtA); free(active); free(excluded);
    return converged ? 0 : 1;
}

/*
 * Lag/shift gather restated in C for the timed CPU baseline of the gather
 * (backend/sglm_pp.py:298-357 shift + :436-457 concat): out[t, c] =
 * X[t - col_shift[c], col_src[c]] when that row exists, else fill.
 */
void sglm_oracle_timeshift(const double *X, long long T, int ldx, const int *col_src,
                           const int *col_shift, int n_out, double fill, double *out) {
    for (long long t = 0; t < T; ++t) {
        double *o = out + (size_t)t * n_out;
        for (int c = 0; c < n_out; ++c) {
            long long ts = t - col_shift[c];
            o[c] = (ts >= 0 && ts < T) ? X[(size_t)ts * ldx + col_src[c]] : fill;
        }
    }
}
