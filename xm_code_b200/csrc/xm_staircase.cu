// xm_staircase.cu — the rank staircase of the reference's entry points behind ONE C-ABI call, xm_solve:
//   XM/src/XM_main.cu:180-310 (solve), :312-401 (solve_rank3), :35-178 (solve_rebuttle).
// Pure host code over the C-ABI calls of this library (xm_trust_region, xm_certify_ex, xm_escape_scale), so it runs on whatever the
// handle holds: dense or block-CSR Q, one GPU or a communicator (every rank makes the same call and gets the same result).
// The pybind11 module (XM/src/XM_main.cpp) and the Python binding (xm_code_b200/solver.py) are thin wrappers around it.
#include "xm_host.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

static void banner(bool on, const char* text) {
    if (!on) return;
    printf("+++++++++++++++++++++++++++++++++\n%s\n+++++++++++++++++++++++++++++++++\n", text);
    fflush(stdout);
}

extern "C" int xm_solve(xm_handle* h, int mode, int max_rank, double tol, double lam, double max_time, const double* s_init,
                        int cert_method, double* R_out, double* s_out, xm_solve_result* res) {
    XmRange nvtx_range("xm_solve");
    if (!h || !R_out || !s_out) return XM_EINVAL;
    if (mode != XM_MODE_FULL && mode != XM_MODE_RANK3 && mode != XM_MODE_REBUTTLE) return XM_EINVAL;
    if (h->N <= 0) { h->err = "no Q set"; return XM_EINVAL; }
    if (mode != XM_MODE_RANK3 && (max_rank < 3 || max_rank > XM_MAX_RANK)) { h->err = "max_rank out of range [3,20]"; return XM_EINVAL; }
    const bool verbose = h->opt.verbose != 0;
    const int n = h->N, n3 = h->n3;
    int o = 3;
    std::vector<double> R0((size_t)n3 * o, 0.0), s0(n, 1.0), v(n3, 0.0);
    if (s_init) std::copy(s_init, s_init + n, s0.begin());      // solve_rebuttle: s_ini is used, R_ini is overwritten by the identity (XM_main.cu:61-63,95-103)
    double gradtol = tol, primal = 0.0;
    int status = 0;
    xm_solve_result out;
    memset(&out, 0, sizeof(out));
    while (o <= max_rank || mode == XM_MODE_RANK3) {
        if (verbose) { printf("+++++++++++++++++++++++++++++++++\nSolve TR with Rank   %d\n+++++++++++++++++++++++++++++++++\n", o); fflush(stdout); }
        std::vector<double> R((size_t)n3 * o), s(n);
        double ls_step = 1.0;
        if (o == 3) {
            std::fill(R0.begin(), R0.end(), 0.0);
            for (int i = 0; i < n; ++i) { R0[3 * i] = 1.0; R0[3 * i + n3 + 1] = 1.0; R0[3 * i + 2 * (size_t)n3 + 2] = 1.0; }
            ls_step = 0.0;
        }
        xm_stats st;
        int rc = xm_trust_region(h, o, R0.data(), s0.data(), lam, &gradtol, ls_step, v.data(), max_time, R.data(), s.data(), &primal, &st, nullptr);
        if (rc) return rc;
        out.n_solves++; out.tcg_iters_total += st.tcg_iters; out.qy_products_total += st.qy_products; out.solve_ms_total += st.solve_ms;
        out.primal = primal; out.gradnorm = st.gradnorm;
        if (mode == XM_MODE_RANK3) { R0 = R; s0 = s; o += 1; break; }
        if (primal < 0) { status = -2; o += 1; break; }                       // line search failed (:244-247): the previous point is kept
        banner(verbose, "Check Eigen value");
        xm_cert_info ci;
        rc = xm_certify_ex(h, o, R.data(), s.data(), lam, primal, cert_method, v.data(), &ci);
        if (rc) return rc;
        out.certified = ci.certified; out.min_eig = ci.min_eig; out.dual = ci.dual; out.gap = ci.gap; out.cert_method = ci.method;
        out.cert_ms_total += ci.ms; out.cert_products_total += ci.products;
        if (ci.certified) {
            o += 1; R0 = R; s0 = s; status = 1;
            break;
        } else if (o < max_rank) {
            R0.assign((size_t)n3 * (o + 1), 0.0);                             // zero-padded new column (:265-269)
            std::copy(R.begin(), R.end(), R0.begin());
            s0 = s;
            xm_escape_scale(n, v.data(), s.data());                           // DecentDirectionKernal (:271)
        } else {
            R0 = R; s0 = s; status = 2;
        }
        o += 1;
    }
    if (verbose && o > max_rank && mode != XM_MODE_RANK3) { printf("BM stoped because max rank\n"); fflush(stdout); }
    const int rank = o - 1;
    std::copy(R0.begin(), R0.begin() + (size_t)n3 * rank, R_out);
    std::copy(s0.begin(), s0.end(), s_out);
    out.rank = rank; out.status = status;
    if (res) *res = out;
    return XM_OK;
}
