/* xm_oracle_c.c — CPU restatement of XMtrustregion in plain C + OpenMP (TEST / BASELINE INFRASTRUCTURE ONLY).
 *
 * The twin of oracle/xm_oracle.py (same control flow, same constants, FP64), built by oracle/Makefile into
 * oracle/_build/libxm_oracle_c.so.  Only tests/ and bench.py's cpu_baseline / --impl reference legs may load it; the
 * shipped product never links or calls it.  It exists because BASELINE.md asks for the CPU baseline to be a compiled
 * restatement running on ALL host cores: the Q.Y product (the only O(N^2) step) is parallel over rows and vectorised over
 * columns, every per-camera loop is parallel over cameras, every reduction is a serial sum of per-camera terms so that
 * results do not depend on the thread count.
 *
 * Reference files restated (paths relative to the reference tree):
 *   XM/include/XM/trustregion.h:77-724   XMtrustregion
 *   XM/include/Dense/batchedQR.h:42-67   modified Gram-Schmidt over the 3 rows of a camera block
 *   XM/include/Dense/matdiagmul.h:28-90  per-camera scaling / per-camera sums (camera 0 has no scale DOF)
 *   XM/include/XM/trustregion.h:18-48    scale kernels (retraction exp, lambda terms)
 * Parity pinning: tests/test_oracle_c.py checks it against the NumPy oracle (itself pinned to outputs of the unmodified
 * reference, tests/test_oracle_vs_reference.py) and directly against the reference-generated goldens tests/golden/ref_*.
 *
 * Layout: a point is Y[(3i+a)*r + j] (camera i's 3 x r block contiguous; the reference's "o x 3N" R_T layout) and s[N]
 * with s[0] == 1 pinned.  Q is 3N x 3N row-major (Q[row*n3 + k]); out = Q X is computed as written (no symmetry assumed).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads only pay off once a sweep is long enough to amortise a fork/join (~10 us) */
static int XMO_PAR = 1;

#define MAX_INNER_ITER 1000 /* trustregion.h:416 */
#define MAX_OUTER_ITER 1000 /* trustregion.h:417 */

typedef struct {
    int outer_iters, tcg_iters, qy_products, status, n_log;
    double primal, gradtol, gradnorm, wall_s;
} xmo_result;

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int xmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

typedef struct {
    int N, r, n3;
    const double* Q;
    double* Xt;   /* operand, j-major: Xt[j*n3 + row] (so that the row sweep vectorises over columns of Q) */
    double* E;    /* result 3N x r (camera-block layout) */
    double* cam;  /* per-camera scratch for deterministic reductions */
} Work;

/* out = alpha * Q X  — Dense/matmul.h:42-87 (cublasDgemm, M = K = 3N, N = r) */
static void qy(const Work* w, const double* X, double alpha, double* out) {
    const int n3 = w->n3, r = w->r;
#pragma omp parallel for schedule(static) if (XMO_PAR)
    for (int row = 0; row < n3; ++row)
        for (int j = 0; j < r; ++j) w->Xt[(size_t)j * n3 + row] = X[(size_t)row * r + j];
#pragma omp parallel for schedule(static) if (XMO_PAR)
    for (int row = 0; row < n3; ++row) {
        const double* q = w->Q + (size_t)row * n3;
        for (int j = 0; j < r; ++j) {
            const double* x = w->Xt + (size_t)j * n3;
            double acc = 0.0;
#pragma omp simd reduction(+ : acc)
            for (int k = 0; k < n3; ++k) acc += q[k] * x[k];
            out[(size_t)row * r + j] = alpha * acc;
        }
    }
}

static double sum_serial(const double* a, int n) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += a[i];
    return t;
}

/* <A, B> over all cameras (serial sum of per-camera dots) */
static double dotR(const Work* w, const double* A, const double* B) {
    const int N = w->N, m = 3 * w->r;
#pragma omp parallel for schedule(static) if (XMO_PAR)
    for (int i = 0; i < N; ++i) {
        double t = 0.0;
        for (int q = 0; q < m; ++q) t += A[(size_t)i * m + q] * B[(size_t)i * m + q];
        w->cam[i] = t;
    }
    return sum_serial(w->cam, N);
}
/* sum_{i>=1} a_i b_i  (the scale part of ProductManifoldInner, trustregion.h:67-74) */
static double dotS(const double* a, const double* b, int N) {
    double t = 0.0;
    for (int i = 1; i < N; ++i) t += a[i] * b[i];
    return t;
}

/* sR = Y * s per camera (matdiagmul.h:28-57) */
static void scale_rows(const Work* w, const double* Y, const double* s, double* out) {
    const int N = w->N, m = 3 * w->r;
#pragma omp parallel for schedule(static) if (XMO_PAR)
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < m; ++q) out[(size_t)i * m + q] = Y[(size_t)i * m + q] * s[i];
}

/* objc (trustregion.h:162-170) given sR: <Q sR, sR> + lam sum_{i>=1} (s_i^2 - 1)^2 ; leaves Q sR in w->E */
static double objective_sR(const Work* w, const double* sR, const double* s, double lam) {
    qy(w, sR, 1.0, w->E);
    double val = dotR(w, w->E, sR);
    double reg = 0.0;
    for (int i = 1; i < w->N; ++i) { const double u = s[i] * s[i] - 1.0; reg += u * u; }
    return val + lam * reg;
}

/* S = sym(A B^T) for one camera: 3 x 3 */
static void sym_outer(const double* A, const double* B, int r, double S[3][3]) {
    double M[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double t = 0.0;
            for (int j = 0; j < r; ++j) t += A[a * r + j] * B[b * r + j];
            M[a][b] = t;
        }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[a][b] = 0.5 * (M[a][b] + M[b][a]);
}

/* batchedQR.h:42-67: normalise row i, then remove it from rows j > i */
static void mgs3(double* A, int r) {
    for (int i = 0; i < 3; ++i) {
        double n = 0.0;
        for (int j = 0; j < r; ++j) n += A[i * r + j] * A[i * r + j];
        n = sqrt(n);
        for (int j = 0; j < r; ++j) A[i * r + j] = A[i * r + j] / n;
        for (int k = i + 1; k < 3; ++k) {
            double d = 0.0;
            for (int j = 0; j < r; ++j) d += A[i * r + j] * A[k * r + j];
            for (int j = 0; j < r; ++j) A[k * r + j] -= d * A[i * r + j];
        }
    }
}

/* log row: k, inner_shown, loss, gradnorm, trstatus, endreason */
int xmo_trust_region(int N, int r, const double* Q, const double* Y0, const double* s0, double lam, double gradtol,
                     double ls_step, const double* v, double max_time, int replicate_stale_sr, double* Y_out, double* s_out,
                     xmo_result* res, double* log6) {
    if (N <= 0 || r < 3 || r > 20 || !Q || !Y0 || !s0 || !Y_out || !s_out || !res || (ls_step != 0.0 && !v)) return -1;
    const double t_start = now_s();
    const int n3 = 3 * N, m = 3 * r;
    XMO_PAR = (N >= 600);
    const size_t VR = (size_t)n3 * r;
    Work w = {N, r, n3, Q, 0, 0, 0};
    /* vectors: Y, Yn, sR, D, G, rg, rR, pR, vR, hvR, X, hr, T(rhr), bestY */
    enum { nR = 14, nS = 12 };
    double* bufR = (double*)calloc(VR * nR + VR /*Xt*/ + VR /*E*/ + (size_t)N, sizeof(double));
    double* bufS = (double*)calloc((size_t)N * nS, sizeof(double));
    double* loss = (double*)calloc(MAX_OUTER_ITER + 2, sizeof(double));
    double* gnorm = (double*)calloc(MAX_OUTER_ITER + 2, sizeof(double));
    if (!bufR || !bufS || !loss || !gnorm) { free(bufR); free(bufS); free(loss); free(gnorm); return -2; }
    double *Y = bufR, *Yn = Y + VR, *sR = Yn + VR, *D = sR + VR, *G = D + VR, *rgR = G + VR, *rR = rgR + VR, *pR = rR + VR,
           *vR = pR + VR, *hvR = vR + VR, *X = hvR + VR, *hr = X + VR, *rhr = hr + VR, *bestY = rhr + VR;
    w.Xt = bestY + VR; w.E = w.Xt + VR; w.cam = w.E + VR;
    double *s = bufS, *sn = s + N, *g = sn + N, *rgs = g + N, *rs = rgs + N, *ps = rs + N, *vs = ps + N, *hvs = vs + N, *hs = hvs + N,
           *rhs = hs + N, *bests = rhs + N, *tmp = bests + N;
    memcpy(Y, Y0, VR * sizeof(double));
    memcpy(s, s0, (size_t)N * sizeof(double));
    s[0] = 1.0;
    const double dim = (double)N * (3.0 * r - 6.0) + (double)N - 1.0; /* :104 */
    const double delta_bar = sqrt(dim);
    double delta = delta_bar / 8.0;
    int nqy = 0, status = 0, n_log = 0;
    memset(res, 0, sizeof(*res));
    scale_rows(&w, Y, s, sR); /* :152 */

    /* ---- rank-escalation line search (:360-408) */
    if (ls_step != 0.0) {
        const double f0 = objective_sR(&w, sR, s, lam); nqy++;
        double alpha = ls_step, f;
        for (int first = 1;; first = 0) {
            if (!first) alpha = alpha / 2;
#pragma omp parallel for schedule(static) if (XMO_PAR)
            for (int i = 0; i < N; ++i) {
                double* A = Yn + (size_t)i * m;
                memcpy(A, Y + (size_t)i * m, (size_t)m * sizeof(double));
                for (int a = 0; a < 3; ++a) A[a * r + (r - 1)] -= alpha * v[3 * i + a]; /* last column = v (:366) */
                mgs3(A, r);
            }
            scale_rows(&w, Yn, s, X);
            f = objective_sR(&w, X, s, lam); nqy++;
            if (!first && alpha < 1e-20) { status = -1; break; } /* :384-391 */
            if (!(f > f0)) break;
        }
        if (status == 0 && f0 - f > 0) { /* :394 */
            memcpy(Y, Yn, VR * sizeof(double));
            if (!replicate_stale_sr) scale_rows(&w, Y, s, sR); /* quirk Q3: the reference keeps the stale sR */
        } else {
            status = -1;
        }
        if (status != 0) {
            res->primal = -1.0; res->status = -1; res->qy_products = nqy; res->gradtol = gradtol;
            memcpy(Y_out, Y, VR * sizeof(double)); memcpy(s_out, s, (size_t)N * sizeof(double));
            free(bufR); free(bufS); free(loss); free(gnorm);
            return 0;
        }
    }
    loss[0] = objective_sR(&w, sR, s, lam); nqy++; /* :422 (with the possibly stale sR) */

    int endreason = 6, trstatus = 4, shrink_count = 0, totalite = 0, i_in = 0, k = 0;
    for (k = 0; k < MAX_OUTER_ITER; ++k) {
        memset(vR, 0, VR * sizeof(double)); memset(hvR, 0, VR * sizeof(double));
        memset(vs, 0, (size_t)N * sizeof(double)); memset(hvs, 0, (size_t)N * sizeof(double));
        memcpy(bestY, Y, VR * sizeof(double)); memcpy(bests, s, (size_t)N * sizeof(double));
        const double bestloss = loss[k];
        /* grad (:186-194) from the stored sR, projection (:307-317), CG initialisation (:476-485) */
        qy(&w, sR, 2.0, D); nqy++;
#pragma omp parallel for schedule(static) if (XMO_PAR)
        for (int i = 0; i < N; ++i) {
            const double* Yi = Y + (size_t)i * m; const double* Di = D + (size_t)i * m;
            double* Gi = G + (size_t)i * m;
            double t = 0.0;
            for (int q = 0; q < m; ++q) { Gi[q] = Di[q] * s[i]; t += Di[q] * Yi[q]; }
            g[i] = (i == 0) ? 0.0 : t + 4.0 * lam * (s[i] * s[i] - 1.0) * s[i];
            double S[3][3];
            sym_outer(Yi, Gi, r, S);
            for (int a = 0; a < 3; ++a)
                for (int j = 0; j < r; ++j) {
                    const double sy = S[a][0] * Yi[j] + S[a][1] * Yi[r + j] + S[a][2] * Yi[2 * r + j];
                    const double rg = Gi[a * r + j] - sy;
                    rgR[(size_t)i * m + a * r + j] = rg; rR[(size_t)i * m + a * r + j] = rg; pR[(size_t)i * m + a * r + j] = -rg;
                }
            rgs[i] = (i == 0) ? 0.0 : g[i] * s[i] * s[i];
            rs[i] = rgs[i]; ps[i] = -rgs[i];
        }
        for (int i = 1; i < N; ++i) tmp[i] = rs[i] / s[i];
        tmp[0] = 0.0;
        double rdotr = dotR(&w, rR, rR) + dotS(tmp, tmp, N); /* :484 */
        gnorm[k] = sqrt(rdotr);
        if (log6 && n_log < MAX_OUTER_ITER + 2) {
            double* L = log6 + 6 * (size_t)n_log;
            L[0] = k; L[1] = i_in + 1; L[2] = loss[k]; L[3] = gnorm[k]; L[4] = (k > 0) ? trstatus : 0; L[5] = (k > 0) ? endreason : 0;
        }
        n_log++;
        if (endreason == 5) break;                                 /* :527 */
        if (gnorm[k] < gradtol) { gradtol /= 10; break; }         /* :532 (quirk Q1) */
        if ((double)(long long)(now_s() - t_start) > max_time) break; /* :538-543 (integer seconds) */
        endreason = 6; trstatus = 4;
        double vdotv = 0.0, vdotp = 0.0, pdotp = rdotr;
        nqy++; /* CsR = 2 Q sR (:553): the reference recomputes the product D already holds */
        for (i_in = 0; i_in < MAX_INNER_ITER; ++i_in) {
            /* ehess (:227-255): X = P s + Y ps ; E = 2 Q X ; then ehess2rhess (:277-295) per camera */
#pragma omp parallel for schedule(static) if (XMO_PAR)
            for (int i = 0; i < N; ++i)
                for (int q = 0; q < m; ++q) X[(size_t)i * m + q] = pR[(size_t)i * m + q] * s[i] + Y[(size_t)i * m + q] * ps[i];
            qy(&w, X, 2.0, w.E); nqy++;
#pragma omp parallel for schedule(static) if (XMO_PAR)
            for (int i = 0; i < N; ++i) {
                const double* Yi = Y + (size_t)i * m; const double* Pi = pR + (size_t)i * m; const double* Di = D + (size_t)i * m;
                const double* Ei = w.E + (size_t)i * m; const double* Gi = G + (size_t)i * m;
                double* hri = hr + (size_t)i * m; double* Ti = rhr + (size_t)i * m;
                double h = 0.0;
                for (int q = 0; q < m; ++q) { hri[q] = Ei[q] * s[i] + Di[q] * ps[i]; h += Ei[q] * Yi[q] + Di[q] * Pi[q]; }
                h += 4.0 * lam * (3.0 * s[i] * s[i] - 1.0) * ps[i];
                hs[i] = (i == 0) ? 0.0 : h;
                double S[3][3], M[3][3];
                sym_outer(Yi, Gi, r, S);
                for (int a = 0; a < 3; ++a)
                    for (int j = 0; j < r; ++j)
                        Ti[a * r + j] = hri[a * r + j] - (S[a][0] * Pi[j] + S[a][1] * Pi[r + j] + S[a][2] * Pi[2 * r + j]);
                sym_outer(Yi, Ti, r, M);
                double out[3 * 20];
                for (int a = 0; a < 3; ++a)
                    for (int j = 0; j < r; ++j)
                        out[a * r + j] = Ti[a * r + j] - (M[a][0] * Yi[j] + M[a][1] * Yi[r + j] + M[a][2] * Yi[2 * r + j]);
                memcpy(Ti, out, (size_t)m * sizeof(double));
                rhs[i] = (i == 0) ? 0.0 : hs[i] * s[i] * s[i] + (ps[i] * s[i]) * g[i];
            }
            for (int i = 1; i < N; ++i) tmp[i] = rhs[i] / (s[i] * s[i]);
            const double alpha = rdotr / (dotR(&w, pR, rhr) + dotS(ps, tmp, N)); /* :566 */
            if (rdotr < 1e-15) { endreason = 5; break; }                          /* :572 */
            if (alpha <= 0 || (vdotv + 2 * alpha * vdotp + alpha * alpha * pdotp > delta * delta)) { /* :577-600 */
                const double tau = (-vdotp + sqrt(vdotp * vdotp + pdotp * (delta * delta - vdotv))) / pdotp;
#pragma omp parallel for schedule(static) if (XMO_PAR)
                for (size_t q = 0; q < VR; ++q) { vR[q] += tau * pR[q]; hvR[q] += tau * rhr[q]; }
                for (int i = 0; i < N; ++i) { vs[i] += tau * ps[i]; hvs[i] += tau * rhs[i]; }
                endreason = (alpha <= 0) ? 1 : 2;
                break;
            }
#pragma omp parallel for schedule(static) if (XMO_PAR)
            for (size_t q = 0; q < VR; ++q) { vR[q] += alpha * pR[q]; rR[q] += alpha * rhr[q]; hvR[q] += alpha * rhr[q]; } /* :605-610 */
            for (int i = 0; i < N; ++i) { vs[i] += alpha * ps[i]; rs[i] += alpha * rhs[i]; hvs[i] += alpha * rhs[i]; }
            for (int i = 1; i < N; ++i) tmp[i] = rs[i] / s[i];
            const double rdotr_new = dotR(&w, rR, rR) + dotS(tmp, tmp, N); /* :626 */
            if (sqrt(rdotr_new) < gnorm[k] * fmin(gnorm[k], 0.1)) { endreason = 3; break; } /* :627 */
            const double beta = rdotr_new / rdotr;
#pragma omp parallel for schedule(static) if (XMO_PAR)
            for (size_t q = 0; q < VR; ++q) pR[q] = beta * pR[q] - rR[q];
            for (int i = 0; i < N; ++i) ps[i] = beta * ps[i] - rs[i];
            const double nvv = vdotv + 2 * alpha * vdotp + alpha * alpha * pdotp; /* :642-644 */
            const double nvp = beta * (vdotp + alpha * pdotp);
            const double npp = beta * beta * pdotp + rdotr_new;
            vdotv = nvv; vdotp = nvp; pdotp = npp;
            rdotr = rdotr_new;
        }
        totalite += i_in + 1; /* :666 (i == max_inner_iter when the loop ran to completion) */
        for (int i = 1; i < N; ++i) tmp[i] = vs[i] / (s[i] * s[i]);
        tmp[0] = 0.0;
        const double loss_qu = (dotR(&w, vR, hvR) + dotS(tmp, hvs, N)) / 2 + (dotR(&w, vR, rgR) + dotS(tmp, rgs, N)); /* :668 */
        if (loss_qu >= 0) break;                                                                                      /* :669 */
        /* retraction (:341-351) */
#pragma omp parallel for schedule(static) if (XMO_PAR)
        for (int i = 0; i < N; ++i) {
            double* A = Y + (size_t)i * m;
            for (int q = 0; q < m; ++q) A[q] += vR[(size_t)i * m + q];
            mgs3(A, r);
            if (i > 0) s[i] = s[i] * exp(vs[i] / s[i]);
        }
        scale_rows(&w, Y, s, sR);                            /* :677 */
        loss[k + 1] = objective_sR(&w, sR, s, lam); nqy++;  /* :678 */
        const double rou = (loss[k + 1] - loss[k]) / loss_qu;
        if (rou < 0.25) { delta = delta * 0.25; trstatus = 1; shrink_count++; }
        else if (rou > 0.75 && endreason <= 2) { delta = fmin(delta * 2, delta_bar); trstatus = 2; shrink_count = 0; }
        else shrink_count = 0;
        if (shrink_count > 3) {
            delta = delta * 1e-3; shrink_count = 0;
            if (delta < 1e-20) break; /* :697-700 */
        }
        if (loss[k + 1] > bestloss || rou < 0.1) { /* :702 reject */
            memcpy(Y, bestY, VR * sizeof(double)); memcpy(s, bests, (size_t)N * sizeof(double));
            loss[k + 1] = bestloss;
            scale_rows(&w, Y, s, sR);
            trstatus = 3;
        }
    }
    if (k > MAX_OUTER_ITER) k = MAX_OUTER_ITER;
    res->primal = loss[k]; /* :715 (k == MAX_OUTER_ITER after a full loop: quirk Q2 reads one past the reference's array) */
    res->gradtol = gradtol; res->outer_iters = k; res->tcg_iters = totalite; res->qy_products = nqy;
    res->gradnorm = gnorm[k < MAX_OUTER_ITER + 2 ? k : MAX_OUTER_ITER + 1]; res->status = 0;
    res->n_log = n_log < MAX_OUTER_ITER + 2 ? n_log : MAX_OUTER_ITER + 2;
    res->wall_s = now_s() - t_start;
    memcpy(Y_out, Y, VR * sizeof(double)); memcpy(s_out, s, (size_t)N * sizeof(double));
    free(bufR); free(bufS); free(loss); free(gnorm);
    return 0;
}
