/*
 * oracle/slm_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the optimisation problems that the reference hands to
 * cvxpy (the arithmetic itself lives in cvxpy + its conic solver, which is not
 * vendored under /root/reference and not installable here; see oracle/README.md).
 * Every convex estimator of the reference reduces to
 *
 *   min_b  1/(2n) ||y - X b||^2  +  sum_j w1_j |b_j|
 *                               +  sum_g w2_g ||b_g||_2
 *                               +  1/2 sum_g delta_g ||b_g||_2^2
 *
 *   data term        src/sparselm/model/_lasso.py:120
 *   l1 term          src/sparselm/model/_lasso.py:107, _adaptive_lasso.py:175,681-683
 *   group term       src/sparselm/model/_lasso.py:254,275,635-637, _adaptive_lasso.py:362,678-680
 *   ridge term       src/sparselm/model/_lasso.py:806-811
 *
 * The solver here is deliberately NOT the algorithm of the CUDA engine: it is a
 * cyclic block-coordinate descent that works on (X, y) directly and keeps the
 * residual vector (no Gram matrix, no momentum), with a per-block majoriser that
 * is verified by backtracking, and it certifies its answer with a duality gap
 * computed from X and y.  For singleton groups it is the same coordinate descent
 * sklearn.linear_model.Lasso uses, which is what the reference's own known-answer
 * test is "borrowed from" (tests/test_lasso.py:30).
 *
 * Layout: X is column-major (n x p, column j at X + j*n); groups are contiguous
 * index ranges gptr[g]..gptr[g+1] (the Python wrapper permutes features).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

static double soft(double v, double t) {
    if (v > t) return v - t;
    if (v < -t) return v + t;
    return 0.0;
}

/* smallest nu >= 0 with || S(g, nu*w1) ||_2 <= nu*w2  (the "epsilon-norm" of the
 * sparse-group penalty); returns +inf when no finite nu exists. The value
 * returned is on the feasible side (>= the exact root). */
static double group_dual_norm(const double *g, const double *w1, i64 s, double w2) {
    double gmax_ratio = 0.0, g2 = 0.0;
    int any_inf = 0;
    for (i64 j = 0; j < s; ++j) {
        double a = fabs(g[j]);
        g2 += a * a;
        if (a > 0.0) {
            if (w1[j] > 0.0) {
                double r = a / w1[j];
                if (r > gmax_ratio) gmax_ratio = r;
            } else {
                any_inf = 1;
            }
        }
    }
    g2 = sqrt(g2);
    if (g2 == 0.0) return 0.0;
    if (w2 <= 0.0) return any_inf ? INFINITY : gmax_ratio;
    /* upper bracket: nu = ||g||/w2 always feasible; nu = gmax_ratio too if finite */
    double hi = g2 / w2;
    if (!any_inf && gmax_ratio < hi) hi = gmax_ratio;
    double lo = 0.0;
    for (int it = 0; it < 200; ++it) {
        double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        double acc = 0.0;
        for (i64 j = 0; j < s; ++j) {
            double u = soft(g[j], mid * w1[j]);
            acc += u * u;
        }
        if (sqrt(acc) <= mid * w2) hi = mid; else lo = mid;
    }
    return hi;
}

/* primal value, duality gap and dual-norm from scratch, all from X and y.
 * out[0]=primal out[1]=dual out[2]=gap out[3]=omega_star */
void slmo_certificate(i64 n, i64 p, const double *X, const double *y, i64 G,
                      const i64 *gptr, const double *w1, const double *w2,
                      const double *delta, const double *beta, double *out) {
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *g = (double *)malloc(sizeof(double) * (size_t)p);
    memcpy(r, y, sizeof(double) * (size_t)n);
    for (i64 j = 0; j < p; ++j) {
        double b = beta[j];
        if (b != 0.0) {
            const double *xj = X + j * n;
            for (i64 i = 0; i < n; ++i) r[i] -= xj[i] * b;
        }
    }
    double rr = 0.0, yr = 0.0;
    for (i64 i = 0; i < n; ++i) { rr += r[i] * r[i]; yr += y[i] * r[i]; }
    double pen = 0.0, ridge = 0.0, omega = 0.0;
    for (i64 gi = 0; gi < G; ++gi) {
        i64 a = gptr[gi], b = gptr[gi + 1];
        double nb = 0.0;
        for (i64 j = a; j < b; ++j) {
            const double *xj = X + j * n;
            double d = 0.0;
            for (i64 i = 0; i < n; ++i) d += xj[i] * r[i];
            g[j] = d / (double)n - delta[gi] * beta[j];
            pen += w1[j] * fabs(beta[j]);
            nb += beta[j] * beta[j];
        }
        pen += w2[gi] * sqrt(nb);
        ridge += delta[gi] * nb;
        double nu = group_dual_norm(g + a, w1 + a, b - a, w2[gi]);
        if (nu > omega) omega = nu;
    }
    double rr_aug = rr + (double)n * ridge;
    double primal = rr_aug / (2.0 * (double)n) + pen;
    double s = (omega > 1.0) ? 1.0 / omega : 1.0;
    if (!(omega < INFINITY)) s = 0.0;
    double dual = (2.0 * s * yr - s * s * rr_aug) / (2.0 * (double)n);
    out[0] = primal; out[1] = dual; out[2] = primal - dual; out[3] = omega;
    free(r); free(g);
}

/* Cyclic block coordinate descent. beta: in = start point, out = solution.
 * info[0]=sweeps done, info[1]=primal, info[2]=gap, info[3]=status (0 gap reached, 1 max sweeps, 2 stationary to rounding)
 * Stops when gap <= tol * max(|primal|, floor_abs). */
int slmo_bcd(i64 n, i64 p, const double *X, const double *y, i64 G, const i64 *gptr,
             const double *w1, const double *w2, const double *delta, double tol,
             double floor_abs, i64 max_sweeps, i64 check_every, double *beta,
             double *info) {
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *L = (double *)malloc(sizeof(double) * (size_t)G);
    i64 smax = 1;
    for (i64 gi = 0; gi < G; ++gi)
        if (gptr[gi + 1] - gptr[gi] > smax) smax = gptr[gi + 1] - gptr[gi];
    double *vnew = (double *)malloc(sizeof(double) * (size_t)smax);
    double *grad = (double *)malloc(sizeof(double) * (size_t)smax);
    double *xd = (double *)malloc(sizeof(double) * (size_t)n);
    double *pv = (double *)malloc(sizeof(double) * (size_t)smax);

    memcpy(r, y, sizeof(double) * (size_t)n);
    for (i64 j = 0; j < p; ++j)
        if (beta[j] != 0.0) {
            const double *xj = X + j * n;
            for (i64 i = 0; i < n; ++i) r[i] -= xj[i] * beta[j];
        }

    /* block curvature: lambda_max(X_g^T X_g)/n estimated by power iteration on X_g
     * (applied as X_g^T (X_g v)); the estimate is only a starting value, every
     * update is checked against the majoriser and L_g is raised if it fails. */
    for (i64 gi = 0; gi < G; ++gi) {
        i64 a = gptr[gi], s = gptr[gi + 1] - a;
        if (s == 1) {
            const double *xj = X + a * n;
            double d = 0.0;
            for (i64 i = 0; i < n; ++i) d += xj[i] * xj[i];
            L[gi] = d / (double)n;
            continue;
        }
        for (i64 k = 0; k < s; ++k) pv[k] = 1.0 / sqrt((double)s) * (1.0 + 0.01 * (double)k);
        double lam = 0.0;
        for (int it = 0; it < 60; ++it) {
            memset(xd, 0, sizeof(double) * (size_t)n);
            for (i64 k = 0; k < s; ++k) {
                const double *xj = X + (a + k) * n;
                double c = pv[k];
                for (i64 i = 0; i < n; ++i) xd[i] += xj[i] * c;
            }
            double nrm = 0.0;
            for (i64 k = 0; k < s; ++k) {
                const double *xj = X + (a + k) * n;
                double d = 0.0;
                for (i64 i = 0; i < n; ++i) d += xj[i] * xd[i];
                vnew[k] = d;
                nrm += d * d;
            }
            nrm = sqrt(nrm);
            if (nrm == 0.0) { lam = 0.0; break; }
            lam = nrm; /* ||A v|| with ||v||=1 */
            for (i64 k = 0; k < s; ++k) pv[k] = vnew[k] / nrm;
        }
        L[gi] = 1.02 * lam / (double)n;
    }

    i64 sweep = 0;
    int status = 1;
    double cert[4] = {0, 0, 0, 0};
    for (sweep = 1; sweep <= max_sweeps; ++sweep) {
        double max_change = 0.0, max_beta = 0.0;
        for (i64 gi = 0; gi < G; ++gi) {
            i64 a = gptr[gi], s = gptr[gi + 1] - a;
            double Lg = L[gi];
            int all_zero = 1;
            for (i64 k = 0; k < s; ++k) {
                const double *xj = X + (a + k) * n;
                double d = 0.0;
                for (i64 i = 0; i < n; ++i) d += xj[i] * r[i];
                grad[k] = -d / (double)n; /* gradient of the LS term */
                if (beta[a + k] != 0.0) all_zero = 0;
            }
            if (Lg <= 0.0) continue; /* null block: columns are all zero */
            for (int bt = 0; bt < 60; ++bt) {
                double un = 0.0;
                for (i64 k = 0; k < s; ++k) {
                    double v = beta[a + k] - grad[k] / Lg;
                    double u = soft(v, w1[a + k] / Lg);
                    vnew[k] = u;
                    un += u * u;
                }
                un = sqrt(un);
                double shrink = 0.0;
                if (un > 0.0) {
                    double t = 1.0 - (w2[gi] / Lg) / un;
                    shrink = (t > 0.0 ? t : 0.0) / (1.0 + delta[gi] / Lg);
                }
                double dn = 0.0;
                for (i64 k = 0; k < s; ++k) {
                    vnew[k] *= shrink;
                    double d = vnew[k] - beta[a + k];
                    dn += d * d;
                }
                if (dn == 0.0) break;
                if (dn > max_change) max_change = dn;
                if (s == 1) { /* exact curvature, no check needed */
                    double d = vnew[0] - beta[a];
                    const double *xj = X + a * n;
                    for (i64 i = 0; i < n; ++i) r[i] -= xj[i] * d;
                    beta[a] = vnew[0];
                    break;
                }
                memset(xd, 0, sizeof(double) * (size_t)n);
                for (i64 k = 0; k < s; ++k) {
                    double d = vnew[k] - beta[a + k];
                    if (d != 0.0) {
                        const double *xj = X + (a + k) * n;
                        for (i64 i = 0; i < n; ++i) xd[i] += xj[i] * d;
                    }
                }
                double q = 0.0;
                for (i64 i = 0; i < n; ++i) q += xd[i] * xd[i];
                if (q / (double)n <= Lg * dn * (1.0 + 1e-12)) {
                    for (i64 i = 0; i < n; ++i) r[i] -= xd[i];
                    for (i64 k = 0; k < s; ++k) beta[a + k] = vnew[k];
                    break;
                }
                Lg *= 1.5; /* majoriser violated: raise curvature and retry */
                L[gi] = Lg;
            }
            (void)all_zero;
        }
        for (i64 j = 0; j < p; ++j)
            if (fabs(beta[j]) > max_beta) max_beta = fabs(beta[j]);
        /* stationary to rounding: a further sweep cannot move the iterate */
        int stationary = sqrt(max_change) <= 4e-16 * max_beta;
        if (stationary || sweep % check_every == 0 || sweep == max_sweeps) {
            slmo_certificate(n, p, X, y, G, gptr, w1, w2, delta, beta, cert);
            double scale = fabs(cert[0]) > floor_abs ? fabs(cert[0]) : floor_abs;
            if (cert[2] <= tol * scale) { status = 0; break; }
            if (stationary) { status = 2; break; }
        }
    }
    if (sweep > max_sweeps) sweep = max_sweeps;
    info[0] = (double)sweep; info[1] = cert[0]; info[2] = cert[2]; info[3] = (double)status;
    free(r); free(L); free(vnew); free(grad); free(xd); free(pv);
    return status;
}

/* Many independent problems on the same X-layout family, one OpenMP thread each:
 * problem k uses X_k = Xs[k] (n_k x p column-major), y_k, its own weights.
 * Used by the CPU baseline (all host cores) -- still the oracle, never the product. */
int slmo_bcd_many(i64 K, const i64 *n, i64 p, const double *const *Xs,
                  const double *const *ys, i64 G, const i64 *gptr, const double *w1,
                  const double *w2, const double *delta, double tol, double floor_rel_y,
                  i64 max_sweeps, i64 check_every, double *betas, double *infos) {
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (i64 k = 0; k < K; ++k) {
        double yy = 0.0;
        for (i64 i = 0; i < n[k]; ++i) yy += ys[k][i] * ys[k][i];
        double floor_abs = floor_rel_y * yy / (2.0 * (double)n[k]);
        bad += slmo_bcd(n[k], p, Xs[k], ys[k], G, gptr, w1 + k * p, w2 + k * G,
                        delta + k * G, tol, floor_abs, max_sweeps, check_every,
                        betas + k * p, infos + k * 4);
    }
    return bad;
}


/* ---- second CPU baseline: Gram-form block coordinate descent along a warm-started path --------
 * What a careful CPU implementation of the same search would do (the strongest CPU arm bench.py
 * reports; still test/bench infrastructure): the training Gram G = X^T X, c = X^T y and y^T y are
 * built once per fold (BLAS, in the caller) and every alpha of a descending path starts from the
 * solution of the previous one.  A block update costs O(p |g|) through q = G beta instead of
 * O(n |g|) through the residual.  Same objective, same duality-gap certificate (evaluated from the
 * Gram quantities), same stopping rule as slmo_bcd.
 * G: p x p row-major (symmetric); w1 [K][p], w2 [K][G], delta [K][G]; betas [K][p] out;
 * infos [K][4] = sweeps, primal, gap, status.  beta0 (may be NULL): start of the first problem. */
static void gram_certificate(i64 p, const double *c, double yty, double n, i64 G, const i64 *gptr,
                             const double *w1, const double *w2, const double *delta,
                             const double *beta, const double *q, double *gbuf, double *out) {
    double cb = 0.0, bq = 0.0, pen = 0.0, ridge = 0.0, omega = 0.0;
    for (i64 gi = 0; gi < G; ++gi) {
        i64 a = gptr[gi], b = gptr[gi + 1];
        double nb = 0.0;
        for (i64 j = a; j < b; ++j) {
            gbuf[j] = (c[j] - q[j]) / n - delta[gi] * beta[j];
            cb += c[j] * beta[j];
            bq += beta[j] * q[j];
            pen += w1[j] * fabs(beta[j]);
            nb += beta[j] * beta[j];
        }
        pen += w2[gi] * sqrt(nb);
        ridge += delta[gi] * nb;
        double nu = group_dual_norm(gbuf + a, w1 + a, b - a, w2[gi]);
        if (nu > omega) omega = nu;
    }
    double rr_aug = yty - 2.0 * cb + bq + n * ridge;
    double yr = yty - cb;
    double primal = rr_aug / (2.0 * n) + pen;
    double s = (omega > 1.0) ? 1.0 / omega : 1.0;
    if (!(omega < INFINITY)) s = 0.0;
    double dual = (2.0 * s * yr - s * s * rr_aug) / (2.0 * n);
    out[0] = primal; out[1] = dual; out[2] = primal - dual; out[3] = omega;
}

int slmo_gram_path(i64 p, const double *Gm, const double *c, double yty, double n, i64 G,
                   const i64 *gptr, i64 K, const double *w1, const double *w2, const double *delta,
                   double tol, double floor_rel_y, i64 max_sweeps, i64 check_every,
                   const double *beta0, double *betas, double *infos) {
    double *beta = (double *)calloc((size_t)p, sizeof(double));
    double *q = (double *)calloc((size_t)p, sizeof(double));
    double *gbuf = (double *)malloc(sizeof(double) * (size_t)p);
    double *L = (double *)malloc(sizeof(double) * (size_t)G);
    i64 smax = 1;
    for (i64 gi = 0; gi < G; ++gi)
        if (gptr[gi + 1] - gptr[gi] > smax) smax = gptr[gi + 1] - gptr[gi];
    double *vnew = (double *)malloc(sizeof(double) * (size_t)smax);
    double *pv = (double *)malloc(sizeof(double) * (size_t)smax);
    double *dl = (double *)malloc(sizeof(double) * (size_t)smax);
    if (beta0) {
        memcpy(beta, beta0, sizeof(double) * (size_t)p);
        for (i64 j = 0; j < p; ++j)
            if (beta[j] != 0.0) {
                const double *gj = Gm + j * p;
                for (i64 i = 0; i < p; ++i) q[i] += gj[i] * beta[j];
            }
    }
    /* block curvature lambda_max(G_gg)/n by power iteration on the diagonal block */
    for (i64 gi = 0; gi < G; ++gi) {
        i64 a = gptr[gi], s = gptr[gi + 1] - a;
        if (s == 1) { L[gi] = Gm[a * p + a] / n; continue; }
        for (i64 k = 0; k < s; ++k) pv[k] = 1.0 / sqrt((double)s) * (1.0 + 0.01 * (double)k);
        double lam = 0.0;
        for (int it = 0; it < 60; ++it) {
            double nrm = 0.0;
            for (i64 k = 0; k < s; ++k) {
                double d = 0.0;
                const double *row = Gm + (a + k) * p + a;
                for (i64 m = 0; m < s; ++m) d += row[m] * pv[m];
                vnew[k] = d;
                nrm += d * d;
            }
            nrm = sqrt(nrm);
            if (nrm == 0.0) { lam = 0.0; break; }
            lam = nrm;
            for (i64 k = 0; k < s; ++k) pv[k] = vnew[k] / nrm;
        }
        L[gi] = 1.02 * lam / n;
    }
    const double floor_abs = floor_rel_y * yty / (2.0 * n);
    int bad = 0;
    for (i64 k = 0; k < K; ++k) {
        const double *w1k = w1 + k * p, *w2k = w2 + k * G, *dk = delta + k * G;
        double cert[4] = {0, 0, 0, 0};
        int status = 1;
        i64 sweep;
        for (sweep = 1; sweep <= max_sweeps; ++sweep) {
            double max_change = 0.0, max_beta = 0.0;
            for (i64 gi = 0; gi < G; ++gi) {
                i64 a = gptr[gi], s = gptr[gi + 1] - a;
                double Lg = L[gi];
                if (Lg <= 0.0) continue;
                for (int bt = 0; bt < 60; ++bt) {
                    double un = 0.0;
                    for (i64 m = 0; m < s; ++m) {
                        double grad = (q[a + m] - c[a + m]) / n;
                        double u = soft(beta[a + m] - grad / Lg, w1k[a + m] / Lg);
                        vnew[m] = u;
                        un += u * u;
                    }
                    un = sqrt(un);
                    double shrink = 0.0;
                    if (un > 0.0) {
                        double t = 1.0 - (w2k[gi] / Lg) / un;
                        shrink = (t > 0.0 ? t : 0.0) / (1.0 + dk[gi] / Lg);
                    }
                    double dn = 0.0;
                    for (i64 m = 0; m < s; ++m) {
                        vnew[m] *= shrink;
                        dl[m] = vnew[m] - beta[a + m];
                        dn += dl[m] * dl[m];
                    }
                    if (dn == 0.0) break;
                    if (s > 1) { /* majoriser check on the diagonal block */
                        double quad = 0.0;
                        for (i64 m = 0; m < s; ++m) {
                            const double *row = Gm + (a + m) * p + a;
                            double d = 0.0;
                            for (i64 t2 = 0; t2 < s; ++t2) d += row[t2] * dl[t2];
                            quad += dl[m] * d;
                        }
                        if (quad / n > Lg * dn * (1.0 + 1e-12)) {
                            Lg *= 1.5;
                            L[gi] = Lg;
                            continue;
                        }
                    }
                    if (dn > max_change) max_change = dn;
                    for (i64 m = 0; m < s; ++m) {
                        if (dl[m] != 0.0) {
                            const double *gj = Gm + (a + m) * p;
                            const double d = dl[m];
                            for (i64 i = 0; i < p; ++i) q[i] += gj[i] * d;
                        }
                        beta[a + m] = vnew[m];
                    }
                    break;
                }
            }
            for (i64 j = 0; j < p; ++j)
                if (fabs(beta[j]) > max_beta) max_beta = fabs(beta[j]);
            int stationary = sqrt(max_change) <= 4e-16 * max_beta;
            if (stationary || sweep % check_every == 0 || sweep == max_sweeps) {
                gram_certificate(p, c, yty, n, G, gptr, w1k, w2k, dk, beta, q, gbuf, cert);
                double scale = fabs(cert[0]) > floor_abs ? fabs(cert[0]) : floor_abs;
                if (cert[2] <= tol * scale) { status = 0; break; }
                if (stationary) { status = 2; break; }
            }
        }
        if (sweep > max_sweeps) sweep = max_sweeps;
        memcpy(betas + k * p, beta, sizeof(double) * (size_t)p);
        infos[k * 4 + 0] = (double)sweep; infos[k * 4 + 1] = cert[0];
        infos[k * 4 + 2] = cert[2]; infos[k * 4 + 3] = (double)status;
        if (status == 1) ++bad;
    }
    free(beta); free(q); free(gbuf); free(L); free(vnew); free(pv); free(dl);
    return bad;
}

/* T independent path segments (fold x alpha range), one OpenMP thread each: segment t works on
 * Gram Gs[t] / cs[t] / yty[t] / ns[t] and the problems koff[t] .. koff[t+1]-1 of the flat
 * w1 / w2 / delta / betas / infos arrays (cold start at the head of every segment). */
int slmo_gram_path_many(i64 T, i64 p, const double *const *Gs, const double *const *cs, const double *yty,
                        const double *ns, i64 G, const i64 *gptr, const i64 *koff, const double *w1,
                        const double *w2, const double *delta, double tol, double floor_rel_y,
                        i64 max_sweeps, i64 check_every, double *betas, double *infos) {
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (i64 t = 0; t < T; ++t) {
        const i64 k0 = koff[t], K = koff[t + 1] - koff[t];
        bad += slmo_gram_path(p, Gs[t], cs[t], yty[t], ns[t], G, gptr, K, w1 + k0 * p, w2 + k0 * G,
                              delta + k0 * G, tol, floor_rel_y, max_sweeps, check_every, NULL,
                              betas + k0 * p, infos + k0 * 4);
    }
    return bad;
}

int slmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
