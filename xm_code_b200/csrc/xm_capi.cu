// xm_capi.cu — host side of libxm_b200.so: the C-ABI declared in include/xm_b200.h.
// There is no CPU fallback anywhere in this file: without an sm_100 device every entry point fails with XM_ENOGPU.
#include "xm_host.h"
#include "xm_solve.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

using namespace xm;

static_assert(sizeof(xm_log_rec) == sizeof(LogRec), "log record layout");
static_assert(XM_LOG_CAP == kLogCap, "log cap");
static_assert(XM_MAX_RANK == kMaxRank, "max rank");

extern "C" void xm_default_options(xm_options* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->device = 0; o->grid_ctas = 0; o->ksplit = 0; o->replicate_stale_sr = 1; o->verbose = 0;
    o->max_outer = 1000; o->max_inner = 1000; o->qy_variant = 0;
}

extern "C" const char* xm_last_error(const xm_handle* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int xm_create(xm_handle** out, const xm_options* opt) {
    if (!out) return XM_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return XM_ENOGPU; }
    xm_handle* h = new (std::nothrow) xm_handle();
    if (!h) return XM_ENOMEM;
    if (opt) h->opt = *opt; else xm_default_options(&h->opt);
    if (h->opt.max_outer <= 0) h->opt.max_outer = 1000;
    if (h->opt.max_inner <= 0) h->opt.max_inner = 1000;
    h->device = h->opt.device;
    if (h->device < 0 || h->device >= ndev) { delete h; return XM_EINVAL; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess || prop.major < 10 || !prop.cooperativeLaunch) {
        delete h; return XM_ENOGPU;      // kernels are built for sm_100a only
    }
    h->num_sm = prop.multiProcessorCount;
    if (cudaSetDevice(h->device) != cudaSuccess) { delete h; return XM_ECUDA; }
    bool ok = cudaMalloc(&h->d_stats, sizeof(DevStats)) == cudaSuccess &&
              cudaMalloc(&h->d_log, sizeof(LogRec) * kLogCap) == cudaSuccess &&
              cudaMalloc(&h->d_bar, 256) == cudaSuccess && cudaMalloc(&h->d_abort, 256) == cudaSuccess &&
              cudaMalloc(&h->d_scalar, 256) == cudaSuccess &&
              cudaMallocHost(&h->h_stats, sizeof(DevStats)) == cudaSuccess &&
              cudaMallocHost(&h->h_log, sizeof(LogRec) * kLogCap) == cudaSuccess;
    if (!ok) { xm_destroy(h); return XM_ENOMEM; }
    *out = h;
    return XM_OK;
}

extern "C" int xm_destroy(xm_handle* h) {
    if (!h) return XM_OK;
    cudaSetDevice(h->device);
    cudaFree(h->Qp); cudaFree(h->Qstage); cudaFree(h->bsr_rowptr); cudaFree(h->bsr_col); cudaFree(h->bsr_val);
    cudaFree(h->ws); cudaFree(h->d_stats); cudaFree(h->d_log); cudaFree(h->d_bar); cudaFree(h->d_abort); cudaFree(h->d_scalar);
    cudaFree(h->io_R0); cudaFree(h->io_s0); cudaFree(h->io_v); cudaFree(h->io_Rout); cudaFree(h->io_sout); cudaFree(h->io_P); cudaFree(h->io_ps);
    cudaFreeHost(h->h_stats); cudaFreeHost(h->h_log);
    delete h;
    return XM_OK;
}

extern "C" int xm_set_stream(xm_handle* h, void* s) {
    if (!h) return XM_EINVAL;
    h->stream = (cudaStream_t)s;
    return XM_OK;
}

// ------------------------------------------------------------------------------------------------ operator upload
static int ensure(xm_handle* h, double** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return XM_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    if (cudaMalloc(p, bytes) != cudaSuccess) { cudaGetLastError(); h->err = "cudaMalloc failed"; return XM_ENOMEM; }
    *cap = bytes;
    return XM_OK;
}

static int set_q_common(xm_handle* h, int n3, const double* q, int64_t ld, bool from_device) {
    if (!h || !q || n3 <= 0 || n3 % 3 != 0 || ld < n3) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    const int ldq = (n3 + 63) / 64 * 64;
    int rc = ensure(h, &h->Qp, &h->Qp_cap, (size_t)n3 * ldq * sizeof(double));
    if (rc) return rc;
    const double* src = q;
    if (!from_device) {
        rc = ensure(h, &h->Qstage, &h->Qstage_cap, (size_t)n3 * n3 * sizeof(double));
        if (rc) return rc;
        XM_CUDA(h, cudaMemcpy2DAsync(h->Qstage, (size_t)n3 * sizeof(double), q, (size_t)ld * sizeof(double),
                                     (size_t)n3 * sizeof(double), n3, cudaMemcpyHostToDevice, h->stream));
        src = h->Qstage; ld = n3;
    }
    dim3 blk(32, 8), grd((ldq + 31) / 32, (n3 + 31) / 32);
    xm_repack_q_kernel<<<grd, blk, 0, h->stream>>>(src, (long long)ld, n3, h->Qp, ldq);
    XM_CUDA(h, cudaGetLastError());
    h->launches++;
    h->n3 = n3; h->N = n3 / 3; h->ldq = ldq; h->is_bsr = false;
    return XM_OK;
}
extern "C" int xm_set_q_dense(xm_handle* h, int n3, const double* q, int64_t ld) { return set_q_common(h, n3, q, ld, false); }
extern "C" int xm_set_q_dense_dev(xm_handle* h, int n3, const double* q, int64_t ld) { return set_q_common(h, n3, q, ld, true); }

extern "C" int xm_set_q_bsr(xm_handle* h, int nb, int bdim, const int* rowptr, const int* colidx, const double* vals) {
    if (!h || !rowptr || !colidx || !vals || nb <= 0 || (bdim != 3 && bdim != 4)) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    const int nnzb = rowptr[nb];
    // re-block on the host into 4x4 row-major padded blocks (128 B each); for bdim==4 only the leading 3x3 acts on
    // the rotation rows (the 4th row/col belongs to translations, which the BM path has already eliminated).
    std::vector<double> blk((size_t)nnzb * 16, 0.0);
    for (int b = 0; b < nnzb; ++b)
        for (int cc = 0; cc < bdim; ++cc)
            for (int rr = 0; rr < bdim; ++rr)
                blk[(size_t)b * 16 + rr * 4 + cc] = vals[(size_t)b * bdim * bdim + (size_t)cc * bdim + rr];
    cudaFree(h->bsr_rowptr); cudaFree(h->bsr_col); cudaFree(h->bsr_val);
    h->bsr_rowptr = nullptr; h->bsr_col = nullptr; h->bsr_val = nullptr;
    XM_CUDA(h, cudaMalloc(&h->bsr_rowptr, sizeof(int) * (nb + 1)));
    XM_CUDA(h, cudaMalloc(&h->bsr_col, sizeof(int) * std::max(nnzb, 1)));
    XM_CUDA(h, cudaMalloc(&h->bsr_val, sizeof(double) * 16 * std::max(nnzb, 1)));
    XM_CUDA(h, cudaMemcpy(h->bsr_rowptr, rowptr, sizeof(int) * (nb + 1), cudaMemcpyHostToDevice));
    XM_CUDA(h, cudaMemcpy(h->bsr_col, colidx, sizeof(int) * nnzb, cudaMemcpyHostToDevice));
    XM_CUDA(h, cudaMemcpy(h->bsr_val, blk.data(), sizeof(double) * 16 * nnzb, cudaMemcpyHostToDevice));
    h->bsr_bdim = bdim; h->is_bsr = true;
    h->N = nb; h->n3 = 3 * nb; h->ldq = (h->n3 + 63) / 64 * 64;
    return XM_OK;
}

// ------------------------------------------------------------------------------------------------ launch planning
struct Plan { int RP, NT, NW, W, cpw, NSW, G, KS, CB; };

static int rank_pad(int r) {
    static const int pads[] = {3, 4, 5, 6, 8, 10, 12, 16, 20};
    for (int p : pads) if (r <= p) return p;
    return -1;
}

static Plan make_plan(const xm_handle* h, int r) {
    Plan p{};
    p.RP = rank_pad(r);
    p.NT = (p.RP <= 10) ? 512 : 256;
    p.NW = p.NT / 32;
    p.W = 4; while (p.W < r) p.W <<= 1;
    p.cpw = 32 / p.W;
    p.NSW = p.NW * p.cpw;
    int G = h->opt.grid_ctas > 0 ? h->opt.grid_ctas : h->num_sm;
    G = std::min(G, h->num_sm);
    G = std::max(1, std::min(G, h->N));
    p.G = G;
    const int cpc = (h->N + G - 1) / G;
    int KS = 1;
    if (h->opt.ksplit > 0) KS = h->opt.ksplit;
    else while (KS * 2 * cpc <= p.NW) KS *= 2;
    KS = std::max(1, std::min(KS, p.NW));
    while (p.NW % KS) KS--;
    if (!h->is_bsr) KS = std::max(1, std::min(KS, h->ldq / 64));
    while (p.NW % KS) KS--;
    p.KS = KS; p.CB = p.NW / KS;
    return p;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve the workspace for (N, r, G); zero it when the geometry changed (operand pad rows must be zero)
static int carve(xm_handle* h, int r, const Plan& p) {
    const size_t N = h->N, n3 = h->n3, ldq = h->ldq;
    const size_t vecR = align_up(n3 * r * sizeof(double), 256), vecS = align_up(N * sizeof(double), 256);
    const size_t total = 11 * vecR + align_up(N * 6 * sizeof(double), 256) + 9 * vecS + align_up((size_t)r * ldq * sizeof(double), 256) +
                         align_up((size_t)kPartialBufs * p.G * kPartialStride * sizeof(double), 256);
    bool fresh = false;
    if (h->ws_cap < total) {
        if (h->ws) cudaFree(h->ws);
        h->ws = nullptr; h->ws_cap = 0;
        if (cudaMalloc(&h->ws, total) != cudaSuccess) { cudaGetLastError(); h->err = "workspace cudaMalloc failed"; return XM_ENOMEM; }
        h->ws_cap = total; fresh = true;
    }
    if (fresh || h->ws_r != r || h->ws_N != (int)N || h->ws_G != p.G || h->ws_ldq != (int)ldq) {
        XM_CUDA(h, cudaMemsetAsync(h->ws, 0, total, h->stream));
        h->ws_r = r; h->ws_N = (int)N; h->ws_G = p.G; h->ws_ldq = (int)ldq;
    }
    char* q = h->ws;
    auto takeR = [&]() { double* t = (double*)q; q += vecR; return t; };
    auto takeS = [&]() { double* t = (double*)q; q += vecS; return t; };
    Dev& d = h->dev;
    d = Dev{};
    d.Y = takeR(); d.Ynew = takeR(); d.D = takeR(); d.Dnew = takeR(); d.EG = takeR(); d.RG = takeR(); d.P = takeR();
    d.Rr = takeR(); d.V = takeR(); d.HV = takeR(); d.HP = takeR();
    d.S6 = (double*)q; q += align_up(N * 6 * sizeof(double), 256);
    d.s = takeS(); d.snew = takeS(); d.gs = takeS(); d.rgs = takeS(); d.ps = takeS(); d.rs = takeS(); d.vs = takeS();
    d.hvs = takeS(); d.hps = takeS();
    d.Xt = (double*)q; q += align_up((size_t)r * ldq * sizeof(double), 256);
    d.partials = (double*)q;
    d.N = (int)N; d.r = r; d.n3 = (int)n3; d.ldq = (int)ldq;
    d.Q = h->is_bsr ? nullptr : h->Qp;
    d.bsr_rowptr = h->bsr_rowptr; d.bsr_col = h->bsr_col; d.bsr_val = h->bsr_val; d.bsr_bdim = h->bsr_bdim;
    d.G = p.G; d.NW = p.NW; d.KS = p.KS; d.CB = p.CB; d.W = p.W; d.cpw = p.cpw; d.NSW = p.NSW;
    d.bar = h->d_bar; d.abort_flag = h->d_abort; d.stats = h->d_stats; d.log = h->d_log;
    d.op_out_scalar = h->d_scalar;
    d.replicate_stale_sr = h->opt.replicate_stale_sr; d.max_outer = h->opt.max_outer; d.max_inner = h->opt.max_inner;
    return XM_OK;
}

static int ensure_io(xm_handle* h, int r) {
    const size_t bR = (size_t)h->n3 * r * sizeof(double), bS = (size_t)h->N * sizeof(double);
    if (h->io_cap_R < bR) {
        double** ps[] = {&h->io_R0, &h->io_Rout, &h->io_P};
        for (auto pp : ps) { if (*pp) cudaFree(*pp); *pp = nullptr; if (cudaMalloc(pp, bR) != cudaSuccess) { cudaGetLastError(); return XM_ENOMEM; } }
        h->io_cap_R = bR;
    }
    if (h->io_cap_s < bS) {
        double** ps[] = {&h->io_s0, &h->io_sout, &h->io_ps};
        for (auto pp : ps) { if (*pp) cudaFree(*pp); *pp = nullptr; if (cudaMalloc(pp, bS) != cudaSuccess) { cudaGetLastError(); return XM_ENOMEM; } }
        if (h->io_v) cudaFree(h->io_v);
        h->io_v = nullptr;
        if (cudaMalloc(&h->io_v, 3 * bS) != cudaSuccess) { cudaGetLastError(); return XM_ENOMEM; }
        h->io_cap_s = bS;
    }
    return XM_OK;
}

template <int RP, int NT>
static cudaError_t launch_solve_t(const Dev& d, cudaStream_t st) {
    void* args[] = {(void*)&d};
    return cudaLaunchCooperativeKernel((const void*)xm_solve_kernel<RP, NT>, dim3(d.G), dim3(NT), args, 0, st);
}
template <int RP, int NT>
static cudaError_t launch_ops_t(const Dev& d, int opcode, cudaStream_t st) {
    void* args[] = {(void*)&d, (void*)&opcode};
    return cudaLaunchCooperativeKernel((const void*)xm_ops_kernel<RP, NT>, dim3(d.G), dim3(NT), args, 0, st);
}
#define XM_DISPATCH(RPV, CALL512, CALL256)                                   \
    switch (RPV) {                                                           \
        case 3:  { constexpr int RP = 3;  constexpr int NT = 512; CALL512; } break;  \
        case 4:  { constexpr int RP = 4;  constexpr int NT = 512; CALL512; } break;  \
        case 5:  { constexpr int RP = 5;  constexpr int NT = 512; CALL512; } break;  \
        case 6:  { constexpr int RP = 6;  constexpr int NT = 512; CALL512; } break;  \
        case 8:  { constexpr int RP = 8;  constexpr int NT = 512; CALL512; } break;  \
        case 10: { constexpr int RP = 10; constexpr int NT = 512; CALL512; } break;  \
        case 12: { constexpr int RP = 12; constexpr int NT = 256; CALL256; } break;  \
        case 16: { constexpr int RP = 16; constexpr int NT = 256; CALL256; } break;  \
        case 20: { constexpr int RP = 20; constexpr int NT = 256; CALL256; } break;  \
        default: e = cudaErrorInvalidValue;                                  \
    }

static cudaError_t launch_solve(const Dev& d, int RPV, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    XM_DISPATCH(RPV, e = (launch_solve_t<RP, NT>(d, st)), e = (launch_solve_t<RP, NT>(d, st)));
    return e;
}
static cudaError_t launch_ops(const Dev& d, int opcode, int RPV, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    XM_DISPATCH(RPV, e = (launch_ops_t<RP, NT>(d, opcode, st)), e = (launch_ops_t<RP, NT>(d, opcode, st)));
    return e;
}

static int prepare(xm_handle* h, int r, Plan* plan) {
    if (!h) return XM_EINVAL;
    if (r < 3 || r > XM_MAX_RANK) { h->err = "rank out of range [3,20]"; return XM_EINVAL; }
    if (h->N <= 0 || (!h->is_bsr && !h->Qp)) { h->err = "no Q set"; return XM_EINVAL; }
    XM_CUDA(h, cudaSetDevice(h->device));
    *plan = make_plan(h, r);
    int rc = carve(h, r, *plan);
    if (rc) return rc;
    rc = ensure_io(h, r);
    if (rc) { h->err = "io staging alloc failed"; return rc; }
    XM_CUDA(h, cudaMemsetAsync(h->d_bar, 0, 256, h->stream));
    XM_CUDA(h, cudaMemsetAsync(h->d_abort, 0, 256, h->stream));
    return XM_OK;
}

static int check_abort(xm_handle* h) {
    int ab = 0;
    XM_CUDA(h, cudaMemcpyAsync(&ab, h->d_abort, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    XM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (ab) { h->err = "device grid barrier timed out"; return XM_ESYNC; }
    return XM_OK;
}

// ------------------------------------------------------------------------------------------------ Q.Y
static int qy_common(xm_handle* h, int r, double alpha, const double* X, double* out, bool dev_ptrs) {
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!X || !out) return XM_EINVAL;
    Dev d = h->dev;
    const cudaMemcpyKind kin = dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // the wire layout (3N x r column-major) IS the operand layout up to the padded leading dimension
    XM_CUDA(h, cudaMemcpy2DAsync(d.Xt, (size_t)d.ldq * sizeof(double), X, (size_t)d.n3 * sizeof(double),
                                 (size_t)d.n3 * sizeof(double), r, kin, h->stream));
    d.qy_alpha = alpha;
    d.op_out_R = dev_ptrs ? out : h->io_Rout;
    XM_CUDA(h, launch_ops(d, 0, p.RP, h->stream));
    h->launches++;
    if (!dev_ptrs) {
        XM_CUDA(h, cudaMemcpyAsync(out, h->io_Rout, (size_t)d.n3 * r * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        XM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return XM_OK;
}
extern "C" int xm_qy(xm_handle* h, int r, double alpha, const double* X, double* out) { return qy_common(h, r, alpha, X, out, false); }
extern "C" int xm_qy_dev(xm_handle* h, int r, double alpha, const double* X, double* out) { return qy_common(h, r, alpha, X, out, true); }

// bench hook: average device time (ms, CUDA events on the handle's stream) of `iters` back-to-back Q.Y launches
// with whatever operand currently sits in the workspace (call xm_qy_dev once first).
extern "C" int xm_bench_qy(xm_handle* h, int r, int iters, double* avg_ms) {
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!avg_ms || iters <= 0) return XM_EINVAL;
    Dev d = h->dev;
    d.qy_alpha = 1.0; d.op_out_R = h->io_Rout;
    cudaEvent_t e0, e1;
    XM_CUDA(h, cudaEventCreate(&e0)); XM_CUDA(h, cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) XM_CUDA(h, launch_ops(d, 0, p.RP, h->stream));
    XM_CUDA(h, cudaEventRecord(e0, h->stream));
    for (int it = 0; it < iters; ++it) XM_CUDA(h, launch_ops(d, 0, p.RP, h->stream));
    XM_CUDA(h, cudaEventRecord(e1, h->stream));
    XM_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    XM_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->launches += iters + 3;
    *avg_ms = (double)ms / iters;
    return XM_OK;
}

// ------------------------------------------------------------------------------------------------ trust region
static void print_log(const DevStats& S, const LogRec* L) {
    // same table as trustregion.h:487-526, printed after the fact (the solve itself never touches the host)
    printf("start linesearch\n");
    for (int n = 0; n < S.n_log; ++n) {
        const LogRec& r = L[n];
        if (r.k > 0) {
            switch (r.trstatus) { case 1: printf("TR- "); break; case 2: printf("TR+ "); break; case 3: printf("REJ "); break; case 4: printf("TR "); break; }
        }
        printf("%d   %d   %1.3e   %1.3e", r.k, r.inner_shown, r.loss, r.gradnorm);
        if (r.k > 0) {
            switch (r.endreason) {
                case 1: printf("   nagative curvature\n"); break;
                case 2: printf("   exceed trust region\n"); break;
                case 3: printf("   reached norm tolerance\n"); break;
                case 5: printf("   numerical issue\n"); break;
                case 6: printf("   max iteration\n"); break;
                default: printf("\n");
            }
        } else printf("\n");
    }
    switch (S.exit_code) {
        case XM_EXIT_RDOTR_TINY: printf("Terminate because of rdotr touched machine precise\n"); break;
        case XM_EXIT_GRADTOL: printf("Terminate because of small gradient norm\n"); break;
        case XM_EXIT_MAXTIME: printf("Terminate because of time limit\n"); break;
        case XM_EXIT_MODEL_INCREASE: printf("error! loss_qu is larger than 0\n"); break;
        case XM_EXIT_DELTA_TINY: printf("delta is too small, BM stopped!\n"); break;
        case XM_EXIT_LINESEARCH_FAILED: printf("linesearch failed! BM stopped! \n"); break;
        default: break;
    }
    printf("\nTotal iteration:     %d\n", S.tcg_iters);
    printf("Time taken by function1: %lld ms\n", (long long)(S.solve_ns / 1000000ull));
    fflush(stdout);
}

static int tr_common(xm_handle* h, int r, const double* R0, const double* s0, double lam, double* gradtol_inout,
                     double ls_step, const double* v, double max_time, double* R_out, double* s_out,
                     double* primal_out, xm_stats* stats, xm_log_rec* log, bool dev_ptrs) {
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!R0 || !s0 || !gradtol_inout || !R_out || !s_out || (ls_step != 0.0 && !v)) { h->err = "null argument"; return XM_EINVAL; }
    Dev d = h->dev;
    const size_t bR = (size_t)d.n3 * r * sizeof(double), bS = (size_t)d.N * sizeof(double);
    if (dev_ptrs) {
        d.R0 = R0; d.s0 = s0; d.vdir = v; d.R_out = R_out; d.s_out = s_out;
    } else {
        XM_CUDA(h, cudaMemcpyAsync(h->io_R0, R0, bR, cudaMemcpyHostToDevice, h->stream));
        XM_CUDA(h, cudaMemcpyAsync(h->io_s0, s0, bS, cudaMemcpyHostToDevice, h->stream));
        if (ls_step != 0.0) XM_CUDA(h, cudaMemcpyAsync(h->io_v, v, 3 * bS, cudaMemcpyHostToDevice, h->stream));
        d.R0 = h->io_R0; d.s0 = h->io_s0; d.vdir = h->io_v; d.R_out = h->io_Rout; d.s_out = h->io_sout;
    }
    d.lam = lam; d.gradtol = *gradtol_inout; d.ls_step = ls_step; d.max_time = max_time;
    cudaEvent_t e0, e1;
    XM_CUDA(h, cudaEventCreate(&e0)); XM_CUDA(h, cudaEventCreate(&e1));
    XM_CUDA(h, cudaEventRecord(e0, h->stream));
    XM_CUDA(h, launch_solve(d, p.RP, h->stream));
    XM_CUDA(h, cudaEventRecord(e1, h->stream));
    h->launches++;
    if (!dev_ptrs) {
        XM_CUDA(h, cudaMemcpyAsync(R_out, h->io_Rout, bR, cudaMemcpyDeviceToHost, h->stream));
        XM_CUDA(h, cudaMemcpyAsync(s_out, h->io_sout, bS, cudaMemcpyDeviceToHost, h->stream));
    }
    XM_CUDA(h, cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, h->stream));
    XM_CUDA(h, cudaMemcpyAsync(h->h_log, h->d_log, sizeof(LogRec) * kLogCap, cudaMemcpyDeviceToHost, h->stream));
    rc = check_abort(h);      // synchronises the stream
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc) return rc;
    const DevStats& S = *h->h_stats;
    if (S.aborted) { h->err = "solver kernel aborted"; return XM_ESYNC; }
    *gradtol_inout = S.gradtol_out;
    if (primal_out) *primal_out = S.primal;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->exit_code = S.exit_code; stats->outer_iters = S.outer_iters; stats->tcg_iters = S.tcg_iters;
        stats->qy_products = S.qy_products; stats->n_log = S.n_log; stats->primal = S.primal; stats->gradnorm = S.gradnorm;
        stats->solve_ms = ms; stats->qy_ms = S.qy_ns * 1e-6; stats->sync_ms = S.sync_ns * 1e-6;
        stats->grid_ctas = p.G; stats->threads_per_cta = p.NT; stats->ksplit = p.KS; stats->launches = 1;
    }
    if (log) memcpy(log, h->h_log, sizeof(LogRec) * (size_t)S.n_log);
    if (h->opt.verbose) print_log(S, h->h_log);
    return XM_OK;
}

extern "C" int xm_trust_region(xm_handle* h, int r, const double* R0, const double* s0, double lam, double* gradtol_inout,
                               double ls_step, const double* v, double max_time, double* R_out, double* s_out,
                               double* primal_out, xm_stats* stats, xm_log_rec* log) {
    return tr_common(h, r, R0, s0, lam, gradtol_inout, ls_step, v, max_time, R_out, s_out, primal_out, stats, log, false);
}
extern "C" int xm_trust_region_dev(xm_handle* h, int r, const double* R0, const double* s0, double lam, double* gradtol_inout,
                                   double ls_step, const double* v, double max_time, double* R_out, double* s_out,
                                   double* primal_out, xm_stats* stats, xm_log_rec* log) {
    return tr_common(h, r, R0, s0, lam, gradtol_inout, ls_step, v, max_time, R_out, s_out, primal_out, stats, log, true);
}

// ------------------------------------------------------------------------------------------------ op-level hooks
static int op_common(xm_handle* h, int r, int opcode, const double* R, const double* s, double lam, const double* P,
                     const double* ps, double lr, double* outR, double* outS, double* out_scalar) {
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!R || !s) return XM_EINVAL;
    Dev d = h->dev;
    const size_t bR = (size_t)d.n3 * r * sizeof(double), bS = (size_t)d.N * sizeof(double);
    XM_CUDA(h, cudaMemcpyAsync(h->io_R0, R, bR, cudaMemcpyHostToDevice, h->stream));
    XM_CUDA(h, cudaMemcpyAsync(h->io_s0, s, bS, cudaMemcpyHostToDevice, h->stream));
    if (P) XM_CUDA(h, cudaMemcpyAsync(h->io_P, P, bR, cudaMemcpyHostToDevice, h->stream));
    if (ps) XM_CUDA(h, cudaMemcpyAsync(h->io_ps, ps, bS, cudaMemcpyHostToDevice, h->stream));
    d.R0 = h->io_R0; d.s0 = h->io_s0; d.op_in_P = h->io_P; d.op_in_ps = h->io_ps; d.op_lr = lr; d.lam = lam;
    d.op_out_R = h->io_Rout; d.op_out_s = h->io_sout;
    XM_CUDA(h, launch_ops(d, opcode, p.RP, h->stream));
    h->launches++;
    if (outR) XM_CUDA(h, cudaMemcpyAsync(outR, h->io_Rout, bR, cudaMemcpyDeviceToHost, h->stream));
    if (outS) XM_CUDA(h, cudaMemcpyAsync(outS, h->io_sout, bS, cudaMemcpyDeviceToHost, h->stream));
    if (out_scalar) XM_CUDA(h, cudaMemcpyAsync(out_scalar, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return check_abort(h);
}
extern "C" int xm_op_objective(xm_handle* h, int r, const double* R, const double* s, double lam, double* f_out) {
    return op_common(h, r, 1, R, s, lam, nullptr, nullptr, 0.0, nullptr, nullptr, f_out);
}
extern "C" int xm_op_rgrad(xm_handle* h, int r, const double* R, const double* s, double lam, double* rgradR, double* rgrads,
                           double* gradnorm_out) {
    return op_common(h, r, 2, R, s, lam, nullptr, nullptr, 0.0, rgradR, rgrads, gradnorm_out);
}
extern "C" int xm_op_rhess(xm_handle* h, int r, const double* R, const double* s, double lam, const double* P, const double* ps,
                           double* HpR, double* Hps) {
    if (!P || !ps) return XM_EINVAL;
    return op_common(h, r, 3, R, s, lam, P, ps, 0.0, HpR, Hps, nullptr);
}
extern "C" int xm_op_retract(xm_handle* h, int r, const double* R, const double* s, const double* etaR, const double* etas,
                             double lr, double* Rn, double* sn) {
    if (!etaR || !etas) return XM_EINVAL;
    return op_common(h, r, 4, R, s, 0.0, etaR, etas, lr, Rn, sn, nullptr);
}

extern "C" int xm_escape_scale(int n, double* v, const double* s) {
    if (!v || !s || n < 0) return XM_EINVAL;
    for (int i = 0; i < n; ++i) { v[3 * i] /= s[i]; v[3 * i + 1] /= s[i]; v[3 * i + 2] /= s[i]; }   // XM_main.cu:8-16
    return XM_OK;
}
