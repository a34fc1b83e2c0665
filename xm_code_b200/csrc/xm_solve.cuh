// xm_solve.cuh — the persistent trust-region kernel and the op-level kernels built from the same phases.
// Reference being replaced: XMtrustregion, XM/include/XM/trustregion.h:77-724 (line numbers cited per phase).
#pragma once
#include "xm_device.cuh"

namespace xm {

struct QMaps { CUtensorMap m[3]; };   // Q tensor maps for the (at most three) distinct batch heights of a launch

// All per-camera phases iterate the CTA's own cameras with the same sub-warp -> camera mapping.
// The loop trip count is warp-uniform; sub-warps past the end run with valid=false on a clamped (in-range) camera
// so that they can take part in the sub-warp shuffles without loading or storing anything.
#define XM_FOR_OWN_CAMERAS(c, i, valid)                                                             \
    for (int _b = (c).cam_lo + (c).warp * (c).cpw; _b < (c).cam_hi; _b += (c).NSW)                  \
        if (const bool valid = (_b + (c).sw < (c).cam_hi); true)                                    \
            if (const int i = valid ? _b + (c).sw : _b; true)

// ---- entry: wire layout (3N x r column-major) -> camera-block state
template <int RP, int NT, bool MG>
__device__ __forceinline__ void phase_load_point(Ctx<RP, NT, MG>& c, const double* R0, const double* s0) {
    const Dev& d = c.d;
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        if (act) {
            double y[3];
            const double* p = R0 + (size_t)c.j * d.n3 + 3 * i;
            y[0] = p[0]; y[1] = p[1]; y[2] = p[2];
            st3(c.R(c.iY), i, d.r, c.j, true, y);
        }
        if (valid && c.j == 0) c.S(c.iS)[i] = (i == 0) ? 1.0 : s0[i];
    }
}
// ---- exit: camera-block state -> wire layout (every rank's output copy gets this CTA's cameras)
template <int RP, int NT, bool MG>
__device__ __forceinline__ void phase_store_point(Ctx<RP, NT, MG>& c, const double* Y, const double* s) {
    const Dev& d = c.d;
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double y[3];
        ld3(Y, i, d.r, c.j, act, y);
        st_out3(c, i, act, y);
        if (valid && c.j == 0) {
            const double si = s[i];
#pragma unroll 1
            for (int w = 0; w < c.world(); ++w) d.outS_peer[w][i] = si;
        }
    }
}

// ---- operand X = s_i * Y_i (for objective / gradient products): trustregion.h:152,375,677
template <int RP, int NT, bool MG>
__device__ __forceinline__ void phase_operand_sR(Ctx<RP, NT, MG>& c, const double* Y, const double* s) {
    const Dev& d = c.d;
    c.begin_push();
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double y[3];
        ld3(Y, i, d.r, c.j, act, y);
        const double si = valid ? s[i] : 0.0;
        double x[3] = {si * y[0], si * y[1], si * y[2]};
        st_operand(c, i, act, x);
    }
}

// ---- line-search trial point: Ynew = MGS(Y - alpha * [0 .. 0 v]) ; operand = s * Ynew   (trustregion.h:360-383)
template <int RP, int NT, bool MG>
__device__ __forceinline__ void phase_ls_trial(Ctx<RP, NT, MG>& c, double alpha) {
    const Dev& d = c.d;
    c.begin_push();
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double a[3];
        ld3(c.R(c.iY), i, d.r, c.j, act, a);
        if (act && c.j == d.r - 1) {
            a[0] -= alpha * d.vdir[3 * i]; a[1] -= alpha * d.vdir[3 * i + 1]; a[2] -= alpha * d.vdir[3 * i + 2];
        }
        if (!act && valid) { a[0] = a[1] = a[2] = 0.0; }
        if (!valid) { a[0] = (c.j == 0); a[1] = (c.j == 1); a[2] = (c.j == 2); }   // keep idle sub-warps finite
        mgs3(a, c.W);
        st3(c.R(c.iYn), i, d.r, c.j, act, a);
        const double si = valid ? c.S(c.iS)[i] : 0.0;
        double x[3] = {si * a[0], si * a[1], si * a[2]};
        st_operand(c, i, act, x);
    }
}

// ---- gradient phase at the current point, from D = 2 Q sR:
//      grad (trustregion.h:186-194) + projection (:307-317) + CG initialisation (:476-485) + first operand (:229-234)
template <int RP, int NT, bool MG>
__device__ __forceinline__ double phase_grad(Ctx<RP, NT, MG>& c, bool build_operand) {
    const Dev& d = c.d;
    const int r = d.r, W = c.W;
    double part = 0.0;
    if (build_operand) c.begin_push();
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double y[3], dd[3];
        ld3(c.R(c.iY), i, r, c.j, act, y); ld3(c.R(c.iD), i, r, c.j, act, dd);
        const double si = valid ? c.S(c.iS)[i] : 1.0;
        double G[3] = {si * dd[0], si * dd[1], si * dd[2]};
        double g = subsum(dd[0] * y[0] + dd[1] * y[1] + dd[2] * y[2], W);
        g += 4.0 * d.lam * (si * si - 1.0) * si;
        if (i == 0) g = 0.0;
        double S[6];
        sym_outer(y, G, W, S);
        double sy[3], rg[3], p[3], z[3] = {0.0, 0.0, 0.0};
        symv(S, y, sy);
#pragma unroll
        for (int a = 0; a < 3; ++a) { rg[a] = G[a] - sy[a]; p[a] = -rg[a]; }
        const double rgs = (i == 0) ? 0.0 : si * si * g;
        st3(c.R(V_EG), i, r, c.j, act, G); st3(c.R(V_RG), i, r, c.j, act, rg); st3(c.R(V_RR), i, r, c.j, act, rg);
        st3(c.R(V_P), i, r, c.j, act, p); st3(c.R(V_V), i, r, c.j, act, z); st3(c.R(V_HV), i, r, c.j, act, z);
        if (valid && c.j == 0) {
            c.S(S_GS)[i] = g; c.S(S_RGS)[i] = rgs; c.S(S_RS)[i] = rgs; c.S(S_PS)[i] = -rgs; c.S(S_VS)[i] = 0.0; c.S(S_HVS)[i] = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) c.s6[(size_t)i * 6 + q] = S[q];
            if (i > 0) { const double t = rgs / si; part += t * t; }
        }
        if (act) part += rg[0] * rg[0] + rg[1] * rg[1] + rg[2] * rg[2];
        if (build_operand) {
            const double psi = -rgs;
            double x[3] = {si * p[0] + psi * y[0], si * p[1] + psi * y[1], si * p[2] + psi * y[2]};
            st_operand(c, i, act, x);
        }
    }
    return part;
}

// ---- tCG update with step alpha (trustregion.h:605-610) and the new residual norm share (:625-626)
template <int RP, int NT, bool MG>
__device__ __forceinline__ double phase_update(Ctx<RP, NT, MG>& c, double alpha) {
    const Dev& d = c.d;
    const int r = d.r;
    double part = 0.0;
    if (d.e_rec) c.begin_push();                // e_rec: this phase also builds the next product's operand X(r_new) = s r_R + r_s Y
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double p[3], hp[3], v[3], rr[3], hv[3];
        ld3(c.R(V_P), i, r, c.j, act, p); ld3(c.R(V_HP), i, r, c.j, act, hp); ld3(c.R(V_V), i, r, c.j, act, v);
        ld3(c.R(V_RR), i, r, c.j, act, rr); ld3(c.R(V_HV), i, r, c.j, act, hv);
#pragma unroll
        for (int a = 0; a < 3; ++a) { v[a] += alpha * p[a]; rr[a] += alpha * hp[a]; hv[a] += alpha * hp[a]; }
        st3(c.R(V_V), i, r, c.j, act, v); st3(c.R(V_RR), i, r, c.j, act, rr); st3(c.R(V_HV), i, r, c.j, act, hv);
        if (act) part += rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2];
        // every lane of the sub-warp computes the new scale residual (lane 0 stores it): the e_rec operand needs it in all lanes
        double rsi = 0.0;
        if (valid && i > 0) rsi = c.S(S_RS)[i] + alpha * c.S(S_HPS)[i];
        __syncwarp();
        if (valid && c.j == 0 && i > 0) {
            const double psi = c.S(S_PS)[i], hpsi = c.S(S_HPS)[i];
            c.S(S_VS)[i] += alpha * psi;
            c.S(S_RS)[i] = rsi;
            c.S(S_HVS)[i] += alpha * hpsi;
            const double t = rsi / c.S(c.iS)[i];
            part += t * t;
        }
        if (d.e_rec) {
            double y[3];
            ld3(c.R(c.iY), i, r, c.j, act, y);
            const double si = valid ? c.S(c.iS)[i] : 0.0;
            const double x[3] = {si * rr[0] + rsi * y[0], si * rr[1] + rsi * y[1], si * rr[2] + rsi * y[2]};
            st_operand(c, i, act, x);
        }
    }
    return part;
}

// ---- boundary / negative-curvature exit: v += tau p ; hv += tau Hp  (trustregion.h:577-600)
template <int RP, int NT, bool MG>
__device__ __forceinline__ void phase_tau(Ctx<RP, NT, MG>& c, double tau) {
    const Dev& d = c.d;
    const int r = d.r;
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double p[3], hp[3], v[3], hv[3];
        ld3(c.R(V_P), i, r, c.j, act, p); ld3(c.R(V_HP), i, r, c.j, act, hp); ld3(c.R(V_V), i, r, c.j, act, v); ld3(c.R(V_HV), i, r, c.j, act, hv);
#pragma unroll
        for (int a = 0; a < 3; ++a) { v[a] += tau * p[a]; hv[a] += tau * hp[a]; }
        st3(c.R(V_V), i, r, c.j, act, v); st3(c.R(V_HV), i, r, c.j, act, hv);
        if (valid && c.j == 0 && i > 0) { c.S(S_VS)[i] += tau * c.S(S_PS)[i]; c.S(S_HVS)[i] += tau * c.S(S_HPS)[i]; }
    }
}

// ---- new search direction p = beta p - r (trustregion.h:634-638) fused with the next operand s*P + ps*Y (:229-234)
template <int RP, int NT, bool MG>
__device__ __forceinline__ void phase_dir(Ctx<RP, NT, MG>& c, double beta) {
    const Dev& d = c.d;
    const int r = d.r;
    c.begin_push();
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double p[3], rr[3], y[3];
        ld3(c.R(V_P), i, r, c.j, act, p); ld3(c.R(V_RR), i, r, c.j, act, rr); ld3(c.R(c.iY), i, r, c.j, act, y);
        const double si = valid ? c.S(c.iS)[i] : 0.0;
        double psi = 0.0;
        if (valid && i > 0) psi = beta * c.S(S_PS)[i] - c.S(S_RS)[i];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = beta * p[a] - rr[a];
        st3(c.R(V_P), i, r, c.j, act, p);
        __syncwarp();                                   // every lane of the sub-warp has read S_PS[i] before lane 0 rewrites it (racecheck)
        if (valid && c.j == 0 && i > 0) c.S(S_PS)[i] = psi;
        double x[3] = {si * p[0] + psi * y[0], si * p[1] + psi * y[1], si * p[2] + psi * y[2]};
        st_operand(c, i, act, x);
    }
}

// ---- model decrease share (trustregion.h:667-668) fused with the retraction (:341-351) and the next operand (:677)
//      etaR = V, etas = vs, lr = 1  (general lr / eta pointers for the op-level hook)
template <int RP, int NT, bool MG>
__device__ __forceinline__ double phase_model_retract(Ctx<RP, NT, MG>& c, const double* etaR, const double* etas, double lr,
                                                      bool with_model) {
    const Dev& d = c.d;
    const int r = d.r;
    double part = 0.0;
    c.begin_push();
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double y[3], v[3];
        ld3(c.R(c.iY), i, r, c.j, act, y); ld3(etaR, i, r, c.j, act, v);
        const double si = valid ? c.S(c.iS)[i] : 1.0;
        const double vsi = (valid && i > 0) ? etas[i] : 0.0;
        if (with_model) {
            double hv[3], rg[3];
            ld3(c.R(V_HV), i, r, c.j, act, hv); ld3(c.R(V_RG), i, r, c.j, act, rg);
            if (act) part += 0.5 * (v[0] * hv[0] + v[1] * hv[1] + v[2] * hv[2]) + (v[0] * rg[0] + v[1] * rg[1] + v[2] * rg[2]);
            if (valid && c.j == 0 && i > 0) {
                const double vsds = vsi / (si * si);
                part += 0.5 * vsds * c.S(S_HVS)[i] + vsds * c.S(S_RGS)[i];
            }
        }
        double a[3] = {y[0] + lr * v[0], y[1] + lr * v[1], y[2] + lr * v[2]};
        if (!act) { a[0] = a[1] = a[2] = 0.0; }
        if (!valid) { a[0] = (c.j == 0); a[1] = (c.j == 1); a[2] = (c.j == 2); }
        mgs3(a, c.W);
        st3(c.R(c.iYn), i, r, c.j, act, a);
        const double sn = (i == 0) ? si : si * exp(lr * vsi / si);   // positiveManifoldRetractionKernal :18-24
        if (valid && c.j == 0) c.S(c.iSn)[i] = sn;
        double x[3] = {sn * a[0], sn * a[1], sn * a[2]};
        st_operand(c, i, act, x);
    }
    return part;
}

#define XM_GSYNC(c) do { if (!(c).grid_sync()) goto xm_abort; } while (0)
// after a phase that built the Q.Y operand and is not followed by a reduction: publish the operand to its consumers
#define XM_OSYNC(c) do { if (!(c).operand_sync()) goto xm_abort; } while (0)
// reduction barrier right after a phase that also built the operand (plain-push protocol: it must publish the remote rows)
#define XM_GSYNC_PUSHED(c) do { if (!(c).grid_sync((c).world() > 1 && (c).d.push_plain != 0)) goto xm_abort; } while (0)

// ================================================================================================ the solver
template <int RP, int NT, int PATH, bool MG>
__global__ void __launch_bounds__(NT, 1) xm_solve_kernel(const __grid_constant__ Dev d, const __grid_constant__ QMaps mapsQ,
                                                         const __grid_constant__ CUtensorMap mapX) {
    __shared__ double red[(NT / 32) * 3 * RP * dense_cams_per_warp(RP, NT)];
    __shared__ double bsum[NT / 32];
    __shared__ double bcast[4];
    extern __shared__ unsigned char dyn_smem[];
    Ctx<RP, NT, MG> c(d, red, bsum, bcast);
    ring_init(c, dyn_smem);
    const bool lead = (blockIdx.x == 0 && c.tid == 0);
    const unsigned long long t_kernel0 = gtimer();
    unsigned long long t_loop0 = t_kernel0;

    // control state — identical in every thread of every CTA (derived only from grid-reduced scalars)
    const int N = d.N, o = d.r;
    const double dim = (double)N * (3.0 * o - 6.0) + (double)N - 1.0;       // :104
    const double delta_bar = sqrt(dim);
    double delta = delta_bar / 8.0;
    double gradtol = d.gradtol;
    int exit_code = 0, nqy = 0, totalite = 0, k = 0, i_inner = 0, n_log = 0;
    int endreason = 6, trstatus = 4, shrink_count = 0;
    bool d_stale = false;
    double loss_k = 0.0, gradnorm = 0.0, f0 = 0.0, fnew = 0.0;

    phase_load_point(c, d.R0, d.s0);
    phase_operand_sR(c, c.R(c.iY), c.S(c.iS));
    XM_OSYNC(c);
    {   // f0 = objc(sR,s) and D = 2 Q sR at the start point (:362 / :422)
        ObjArgs oa{c.R(c.iY), c.S(c.iS), c.R(c.iD)};
        double part = qy_phase<RP, NT, MODE_OBJ, PATH>(c, oa, mapsQ.m, &mapX);
        c.publish(part); XM_GSYNC(c); f0 = c.collect(); nqy++;
    }
    loss_k = f0;
    if (d.ls_step != 0.0) {          // rank-escalation line search (:360-408)
        double alpha = d.ls_step;
        bool failed = false, first = true;
        for (;;) {
            if (!first) alpha = alpha / 2;                                   // :378
            phase_ls_trial(c, alpha);
            XM_OSYNC(c);
            ObjArgs oa{c.R(c.iYn), c.S(c.iS), c.R(c.iDn)};
            double part = qy_phase<RP, NT, MODE_OBJ, PATH>(c, oa, mapsQ.m, &mapX);
            c.publish(part); XM_GSYNC(c); fnew = c.collect(); nqy++;
            if (!first && alpha < 1e-20) { failed = true; break; }            // :384-391 (tested after the evaluation)
            if (!(fnew > f0)) break;                                         // while (f > f0)
            first = false;
        }
        if (!failed && (f0 - fnew > 0)) {                                    // :394
            int t = c.iY; c.iY = c.iYn; c.iYn = t;                      // R_T, R <- new (:396-397)
            if (d.replicate_stale_sr) { d_stale = true; }                    // quirk Q3: sR, loss[0], D stay at the old point
            else { t = c.iD; c.iD = c.iDn; c.iDn = t; loss_k = fnew; }
        } else {
            exit_code = -1;                                                   // line search failed: primal = -1 (:384-405)
            loss_k = -1.0;
            goto xm_finish;
        }
    }
    t_loop0 = gtimer();

    for (k = 0; k < d.max_outer; ++k) {
        double tflag = 0.0;
        if (c.gc == 0 && c.tid == 0) tflag = ((double)((gtimer() - t_loop0) / 1000000000ull) > d.max_time) ? 1.0 : 0.0;   // :538-543 (one clock decides for all ranks)
        const double part = phase_grad(c, true);
        c.unpack_operand();                      // multi-GPU: the peers' operand rows, before the reduction barrier publishes them
        c.publish(part, tflag); XM_GSYNC_PUSHED(c);
        double timeflag = 0.0;
        double rdotr = c.collect(&timeflag);
        gradnorm = sqrt(rdotr);
        if (lead && n_log < kLogCap) {
            LogRec& L = d.log[n_log];
            L.k = k; L.inner_shown = i_inner + 1; L.trstatus = (k > 0) ? trstatus : 0; L.endreason = (k > 0) ? endreason : 0;
            L.loss = loss_k; L.gradnorm = gradnorm; L.delta = delta;
        }
        n_log++;
        if (endreason == 5) { exit_code = 2; break; }                               // :527-530
        if (gradnorm < gradtol) { gradtol /= 10; exit_code = 1; break; }            // :532-536
        if (timeflag != 0.0) { exit_code = 3; break; }
        endreason = 6; trstatus = 4;
        double vdotv = 0.0, vdotp = 0.0, pdotp = rdotr;

        c.erec_first = true;
        for (i_inner = 0; i_inner < d.max_inner; ++i_inner) {                       // :559-664
            ObjArgs oa{nullptr, nullptr, nullptr};
            c.trace_on = (d.profile && lead && nqy >= 200 && nqy < 202);
            c.tr(1);
            const double ph = qy_phase<RP, NT, MODE_HESS, PATH>(c, oa, mapsQ.m, &mapX);
            c.erec_first = false;
            c.tr(2);
            c.publish(ph); XM_GSYNC(c);
            c.tr(3);
            const double pHp = c.collect(); nqy++;
            const double alpha = rdotr / pHp;                                       // :566
            if (rdotr < 1e-15) { endreason = 5; break; }                            // :572-576
            const bool exceed = (vdotv + 2 * alpha * vdotp + alpha * alpha * pdotp > delta * delta);
            if (alpha <= 0 || exceed) {                                             // :577-600
                const double sq = sqrt(vdotp * vdotp + pdotp * (delta * delta - vdotv));
                const double tau = (-vdotp + sq) / pdotp;
                phase_tau(c, tau);
                endreason = (alpha <= 0) ? 1 : 2;
                break;
            }
            c.tr(4);
            const double pr = phase_update(c, alpha);
            c.tr(5);
            if (d.e_rec) {                          // the operand X(r_new) rides on the <r,r> reduction barrier: two barriers per iteration
                c.unpack_operand();
                c.publish(pr); XM_GSYNC_PUSHED(c);
            } else {
                c.publish(pr); XM_GSYNC(c);
            }
            c.tr(6);
            const double rdotr_new = c.collect();                                   // :626
            if (sqrt(rdotr_new) < gradnorm * fmin(gradnorm, 0.1)) { endreason = 3; break; }   // :627-630
            const double beta = rdotr_new / rdotr;
            c.tr(7);
            if (d.e_rec) {
                c.erec_beta = beta;                 // p = beta p - r and E = beta E - 2 Q X(r) happen in the next product's epilogue
            } else {
                phase_dir(c, beta);
                c.tr(8);
                XM_OSYNC(c);
            }
            c.tr(9);
            const double nvv = vdotv + 2 * alpha * vdotp + alpha * alpha * pdotp;   // :642-644
            const double nvp = beta * (vdotp + alpha * pdotp);
            const double npp = beta * beta * pdotp + rdotr_new;
            vdotv = nvv; vdotp = nvp; pdotp = npp;
            rdotr = rdotr_new;
        }
        totalite += i_inner + 1;                                                    // :666

        const double pm = phase_model_retract(c, c.R(V_V), c.S(S_VS), 1.0, true);
        c.unpack_operand();
        c.publish(pm); XM_GSYNC_PUSHED(c);
        const double loss_qu = c.collect();
        if (loss_qu >= 0) { exit_code = 4; break; }                                 // :669-672
        {
            ObjArgs oa{c.R(c.iYn), c.S(c.iSn), c.R(c.iDn)};
            const double pf = qy_phase<RP, NT, MODE_OBJ, PATH>(c, oa, mapsQ.m, &mapX);
            c.publish(pf); XM_GSYNC(c); fnew = c.collect(); nqy++;                 // loss[k+1] (:678)
        }
        const double rou = (fnew - loss_k) / loss_qu;                               // :680
        if (rou < 0.25) { delta = delta * 0.25; trstatus = 1; shrink_count++; }
        else if (rou > 0.75 && endreason <= 2) { delta = fmin(delta * 2, delta_bar); trstatus = 2; shrink_count = 0; }
        else { shrink_count = 0; }
        if (shrink_count > 3) {
            delta = delta * 1e-3; shrink_count = 0;
            if (delta < 1e-20) {                                                    // :697-700
                // the reference breaks here with R,s already overwritten by the new point and primal = loss[k]
                int t = c.iY; c.iY = c.iYn; c.iYn = t; t = c.iS; c.iS = c.iSn; c.iSn = t;
                exit_code = 5; break;
            }
        }
        if ((fnew > loss_k) || (rou < 0.1)) {                                       // :702 reject (bestloss == loss[k])
            trstatus = 3;                                                           // loss[k+1] = bestloss
            if (d_stale) {          // the reference recomputes everything from the restored (fresh) sR next iteration
                phase_operand_sR(c, c.R(c.iY), c.S(c.iS));
                XM_OSYNC(c);
                ObjArgs oa{c.R(c.iY), c.S(c.iS), c.R(c.iD)};
                (void)qy_phase<RP, NT, MODE_OBJ, PATH>(c, oa, mapsQ.m, &mapX); nqy++;
                XM_GSYNC(c);        // nobody may rewrite the operand while another CTA is still sweeping it
            }
        } else {
            int t = c.iY; c.iY = c.iYn; c.iYn = t; t = c.iS; c.iS = c.iSn; c.iSn = t;
            t = c.iD; c.iD = c.iDn; c.iDn = t;
            loss_k = fnew;
        }
        d_stale = false;
    }
    if (k >= d.max_outer && exit_code == 0) exit_code = 6;

xm_finish:
    ring_drain(c);
    phase_store_point(c, c.R(c.iY), c.S(c.iS));
    // multi-GPU: the result rows were pushed into every rank's output copy; nobody's host may read its copy (or launch the
    // next solve, which rewrites the peers' operand) before all of them have landed
    if (c.world() > 1) { if (!c.grid_sync(true)) goto xm_abort; }
    c.save_epoch();
    if (lead) {
        DevStats& S = *d.stats;
        S.exit_code = exit_code; S.outer_iters = k; S.tcg_iters = totalite; S.qy_products = nqy;
        S.n_log = n_log < kLogCap ? n_log : kLogCap; S.aborted = 0;
        S.primal = loss_k; S.gradnorm = gradnorm; S.gradtol_out = gradtol;
        S.solve_ns = gtimer() - t_kernel0; S.qy_ns = c.t_qy; S.sync_ns = c.t_sync;
        S.dbg[0] = c.dbg0; S.dbg[1] = c.dbg1; S.dbg[2] = c.dbg2; S.dbg[3] = c.dbg3;
    }
    return;
xm_abort:
    ring_drain(c);
    if (lead) { d.stats->aborted = 1; d.stats->exit_code = 0; }
    return;
}

// ================================================================================================ op-level kernels
// opcode: 0 = Q.Y (operand already in Xt)                          -> op_out_R (3N x r col-major) = qy_alpha * Q X
//         1 = objective at (R0,s0)                                  -> op_out_scalar[0]
//         2 = Riemannian gradient at (R0,s0)                        -> op_out_R = rgradR, op_out_s = rgrads, scalar[0] = gradnorm
//         3 = Riemannian Hessian-vector at (R0,s0) along (P,ps)     -> op_out_R = HpR, op_out_s = Hps
//         4 = retraction of (R0,s0) along (P,ps) with step op_lr    -> op_out_R = Rn, op_out_s = sn
template <int RP, int NT, int PATH, bool MG>
__device__ __forceinline__ bool ops_body(Ctx<RP, NT, MG>& c, const Dev& d, const QMaps& mapsQ, const CUtensorMap& mapX, const int opcode) {
    if (opcode == 0) {
        ObjArgs oa{nullptr, nullptr, nullptr};
        // op_repeat > 1 (xm_bench_qy): back-to-back products inside one launch, ring prefetch across them as in the solver
        const int nrep = d.op_repeat < 0 ? -d.op_repeat : d.op_repeat;
        for (int rep = 0; rep < nrep; ++rep) {
            (void)qy_phase<RP, NT, MODE_OUT, PATH>(c, oa, mapsQ.m, &mapX, rep + 1 < nrep);
            if (d.op_repeat < 0) { XM_GSYNC(c); }      // lock-step products, as inside the solver
            else __syncthreads();
        }
        if (blockIdx.x == 0 && c.tid == 0) { d.stats->dbg[0] = c.dbg0; d.stats->dbg[1] = c.dbg1; d.stats->dbg[2] = c.dbg2; d.stats->dbg[3] = c.dbg3; d.stats->qy_ns = c.t_qy; d.stats->sync_ns = c.t_sync; }
        return true;
    }
    if (opcode == 5) {                 // grid-barrier micro-benchmark: |op_repeat| barriers back to back
        const int nrep = d.op_repeat < 0 ? -d.op_repeat : d.op_repeat;
        for (int rep = 0; rep < nrep; ++rep) { XM_GSYNC(c); }
        if (c.tid == 0 && blockIdx.x == 0) for (int q = 0; q < 5; ++q) d.stats->dbg[q] = c.bseg[q];          // leader's segments
        if (c.tid == 0 && blockIdx.x == 1) { d.stats->dbg[5] = c.bseg[0]; d.stats->dbg[6] = c.bseg[3]; d.stats->dbg[7] = c.bseg[4]; }
        if (d.op_repeat < 0) {            // negative: as many operand exchanges (push + unpack + local barrier) on top, timed by the caller
            for (int rep = 0; rep < nrep; ++rep) {
                c.begin_push();
                XM_FOR_OWN_CAMERAS(c, i, valid) { const double x[3] = {1.0 * rep, 2.0, 3.0}; st_operand(c, i, c.act && valid, x); }
                XM_OSYNC(c);
            }
        }
        return true;
    }
    if (opcode == 6) {                 // the 3x3 diagonal blocks of the operator, 9 doubles per camera (row-major), into every rank's
        XM_FOR_OWN_CAMERAS(c, i, valid) {   // output copy — the block-Jacobi preconditioner of the iterative certificate needs all of them
            if (valid && c.j < 3) {
                const int a = c.j;
                double q[3] = {0.0, 0.0, 0.0};
                if (PATH == 2) {
                    for (int b = d.bsr_rowptr[i - d.cam0]; b < d.bsr_rowptr[i - d.cam0 + 1]; ++b)
                        if (d.bsr_col[b] == i) { const double* blk = d.bsr_val + (size_t)b * 16 + 4 * a; q[0] = blk[0]; q[1] = blk[1]; q[2] = blk[2]; }
                } else {
                    const double* row = d.Q + (size_t)(3 * i + a - d.row0) * d.ldq + 3 * i;
                    q[0] = row[0]; q[1] = row[1]; q[2] = row[2];
                }
#pragma unroll 1
                for (int w = 0; w < c.world(); ++w) { double* o = d.outR_peer[w] + (size_t)i * 9 + 3 * a; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
            }
        }
        return true;
    }
    phase_load_point(c, d.R0, d.s0);
    if (opcode == 4) {
        // direction arrives in wire layout: convert into V / vs first
        XM_FOR_OWN_CAMERAS(c, i, valid) {
            const bool act = c.act && valid;
            if (act) {
                const double* p = d.op_in_P + (size_t)c.j * d.n3 + 3 * i;
                double v[3] = {p[0], p[1], p[2]};
                st3(c.R(V_V), i, d.r, c.j, true, v);
            }
            if (valid && c.j == 0) c.S(S_VS)[i] = (i == 0) ? 0.0 : d.op_in_ps[i];
        }
        (void)phase_model_retract(c, c.R(V_V), c.S(S_VS), d.op_lr, false);
        phase_store_point(c, c.R(c.iYn), c.S(c.iSn));
        return true;
    }
    phase_operand_sR(c, c.R(c.iY), c.S(c.iS));
    XM_OSYNC(c);
    {
        ObjArgs oa{c.R(c.iY), c.S(c.iS), c.R(c.iD)};
        double part = qy_phase<RP, NT, MODE_OBJ, PATH>(c, oa, mapsQ.m, &mapX, opcode == 3);
        c.publish(part); XM_GSYNC(c);
        const double f = c.collect();
        if (opcode == 1) { if (blockIdx.x == 0 && c.tid == 0) d.op_out_scalar[0] = f; return true; }
    }
    {
        double part = phase_grad(c, false);
        c.publish(part); XM_GSYNC(c);
        const double rd = c.collect();
        if (opcode == 2) {
            if (blockIdx.x == 0 && c.tid == 0) d.op_out_scalar[0] = sqrt(rd);
            phase_store_point(c, c.R(V_RG), c.S(S_RGS));
            return true;
        }
    }
    // opcode 3: overwrite P / ps with the caller's direction, build the operand, one Hessian-vector product
    c.begin_push();
    __syncthreads();
    XM_FOR_OWN_CAMERAS(c, i, valid) {
        const bool act = c.act && valid;
        double p[3] = {0, 0, 0}, y[3];
        ld3(c.R(c.iY), i, d.r, c.j, act, y);
        if (act) {
            const double* pp = d.op_in_P + (size_t)c.j * d.n3 + 3 * i;
            p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2];
        }
        st3(c.R(V_P), i, d.r, c.j, act, p);
        const double si = valid ? c.S(c.iS)[i] : 0.0;
        const double psi = (valid && i > 0) ? d.op_in_ps[i] : 0.0;
        if (valid && c.j == 0) c.S(S_PS)[i] = psi;
        double x[3] = {si * p[0] + psi * y[0], si * p[1] + psi * y[1], si * p[2] + psi * y[2]};
        st_operand(c, i, act, x);
    }
    XM_OSYNC(c);
    {
        ObjArgs oa{nullptr, nullptr, nullptr};
        (void)qy_phase<RP, NT, MODE_HESS, PATH>(c, oa, mapsQ.m, &mapX, false);
        phase_store_point(c, c.R(V_HP), c.S(S_HPS));
    }
    return true;
xm_abort:
    return false;
}

template <int RP, int NT, int PATH, bool MG>
__global__ void __launch_bounds__(NT, 1) xm_ops_kernel(const __grid_constant__ Dev d, const __grid_constant__ QMaps mapsQ,
                                                       const __grid_constant__ CUtensorMap mapX, const int opcode) {
    __shared__ double red[(NT / 32) * 3 * RP * dense_cams_per_warp(RP, NT)];
    __shared__ double bsum[NT / 32];
    __shared__ double bcast[4];
    extern __shared__ unsigned char dyn_smem[];
    Ctx<RP, NT, MG> c(d, red, bsum, bcast);
    ring_init(c, dyn_smem);
    bool ok = ops_body<RP, NT, PATH>(c, d, mapsQ, mapX, opcode);
    // multi-GPU: results were pushed to every rank and the peers' operand copies may still be in use — leave together
    if (ok && c.world() > 1) ok = c.grid_sync(true);
    if (ok) c.save_epoch();
    else ring_drain(c);
}

}  // namespace xm
