/* xm_b200.h — C-ABI of libxm_b200.so: the B200-native (sm_100a) replacement for the hot path of XM's
 * Burer-Monteiro solver.  Plain C types only; every buffer is caller-owned; no exceptions cross this boundary.
 *
 * The reference has no FFI of its own for this path: its pybind11 functions (XM/src/XM_main.cu:403-408) call the
 * header-only C++ XMtrustregion / checkeig directly.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference tree):
 *
 *   xm_set_q_dense      <- loadCMatrixFromBin + opt_var C({3n,3n}); C.SynchronizeHostToDevice   XM/src/XM_main.cu:18-33,189-192
 *   xm_qy               <- DnMatDnMat (cublasDgemm, M=K=3N, N=r)                                XM/include/Dense/matmul.h:42-87
 *   xm_trust_region     <- XMtrustregion(C,R0,s0,R,s,lam,gradtol&,ls_step,v,&primal,maxtime)    XM/include/XM/trustregion.h:77-724
 *   xm_op_*             <- the lambdas objc/grad/projection/ehess+ehess2rhess/retraction        XM/include/XM/trustregion.h:162-351
 *   xm_certify(_ex)     <- checkeig(C,sR,lam,v,primal)                                          XM/include/XM/checkeig.h:42-368
 *   xm_solve            <- solve / solve_rank3 / solve_rebuttle (the rank staircase)           XM/src/XM_main.cu:35-401
 *   xm_escape_scale     <- DecentDirectionKernal                                                XM/src/XM_main.cu:8-16
 *   xm_create_matrix    <- create_matrix(weight, edges, landmarks, output_path)                 utils/creatematrix.py:52-341
 *   xm_recover          <- recover_XM(Q,R,s,Abar,lam)                                           utils/recoversolution.py:4-86
 *   xm_residuals        <- the per-observation error of the XM^2 outlier cut                    3_test_colmap_glomap.py:304-316
 *
 * Matrix arguments use the reference's wire layouts: R is 3N x r COLUMN-MAJOR with
 * rows 3i..3i+2 = camera i; s is length N with s[0] == 1 (the reference's s_ex); v is length 3N.
 * Host pointers unless the function name ends in _dev (then: device pointers valid on the handle's device).
 * Return value: XM_OK (0) or a negative XM_E* code.  Numerical exits of the algorithm are NOT errors; they are
 * reported in xm_stats.exit_code (mirrors the reference's explicit exits, trustregion.h:384-405,527-543,669-700).
 * One handle = one CUDA device; a handle is not thread-safe.
 */
#ifndef XM_B200_H
#define XM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xm_handle xm_handle;

enum {
    XM_OK = 0,
    XM_EINVAL = -1,      /* bad argument (null pointer, r out of [3,XM_MAX_RANK], no Q set, ...) */
    XM_ECUDA = -2,       /* a CUDA runtime call failed; xm_last_error() has the text */
    XM_ENOMEM = -3,
    XM_ENOGPU = -4,      /* no sm_100 device / kernels cannot launch: there is NO CPU fallback */
    XM_ESYNC = -5,       /* device-side grid barrier timed out (kernel aborted itself) */
    XM_EUNSUPPORTED = -6
};

#define XM_MAX_RANK 20

/* xm_stats.exit_code: why xm_trust_region returned */
enum {
    XM_EXIT_GRADTOL = 1,      /* gradnorm < gradtol (gradtol_inout is divided by 10, quirk Q1, trustregion.h:532-535) */
    XM_EXIT_RDOTR_TINY = 2,   /* tCG residual below 1e-15 ("numerical issue", endreason 5, :527-530) */
    XM_EXIT_MAXTIME = 3,      /* :538-543 */
    XM_EXIT_MODEL_INCREASE = 4, /* loss_qu >= 0 (:669-672) */
    XM_EXIT_DELTA_TINY = 5,   /* trust radius < 1e-20 (:697-700) */
    XM_EXIT_MAX_OUTER = 6,    /* 1000 outer iterations */
    XM_EXIT_LINESEARCH_FAILED = -1 /* rank-escalation line search failed; primal_out = -1 (:384-405) */
};

typedef struct xm_options {
    int device;             /* CUDA device ordinal */
    int grid_ctas;          /* 0 = auto (<= one CTA per SM); tuning/test hook */
    int ksplit;             /* 0 = auto; number of k-splits of a camera's Q rows inside a CTA (1,2,4,8,16) */
    int replicate_stale_sr; /* 1 (default) = reproduce reference quirk Q3 after an accepted line search */
    int verbose;            /* 1 = print the reference's per-outer-iteration table to stdout after the solve */
    int max_outer;          /* default 1000 (trustregion.h:417) */
    int max_inner;          /* default 1000 (trustregion.h:416) */
    int qy_variant;         /* dense Q.Y path: 0 = auto (2-D TMA ring + cross-phase prefetch), 1 = direct streaming loads */
    int vec_in_global;      /* 1 = keep the per-camera state vectors in HBM/L2 even when they would fit in shared memory */
    int profile;            /* 1 = fine-grained in-kernel phase timers (xm_stats.phase_ms); costs a few percent */
    int three_barrier_tcg;  /* 0 (default) = two grid barriers per tCG iteration: E = 2 Q X(p) is kept by the recurrence
                               E <- beta E - 2 Q X(r_new), so the product's operand is built from the new residual and rides on the
                               <r,r> reduction barrier (same iterates up to rounding; measured +5 % on the latency-bound config);
                               1 = the reference's order of operations: operand from the new direction, three barriers */
} xm_options;

typedef struct xm_log_rec { /* one line of the reference's stdout table (trustregion.h:487-526) */
    int k, inner_shown, trstatus, endreason;
    double loss, gradnorm, delta;
} xm_log_rec;

typedef struct xm_stats {
    int exit_code;
    int outer_iters;        /* value of k at exit */
    int tcg_iters;          /* the reference's "Total iteration" (sum over outer iterations of i+1) */
    int qy_products;        /* Q.Y products actually executed on the device */
    int n_log;              /* number of valid entries in log (<= XM_LOG_CAP) */
    double primal;          /* loss[k] */
    double gradnorm;        /* last Riemannian gradient norm */
    double solve_ms;        /* device time of the persistent solve kernel (CUDA events) */
    double qy_ms;           /* device time spent inside Q.Y sweeps (globaltimer, CTA 0) */
    double sync_ms;         /* device time CTA 0 spent waiting in grid barriers */
    int grid_ctas, threads_per_cta, ksplit, launches; /* launch configuration actually used / kernels launched */
    double phase_ms[4];     /* CTA 0 breakdown of the Q.Y phases: first-tile wait, later tile waits, tile math, reduce+epilogue */
} xm_stats;

#define XM_LOG_CAP 1002

void xm_default_options(xm_options* opt);
int  xm_create(xm_handle** out, const xm_options* opt);
int  xm_destroy(xm_handle* h);
const char* xm_last_error(const xm_handle* h);
/* stream: a cudaStream_t (as void*) on the handle's device; NULL = the legacy default stream. */
int  xm_set_stream(xm_handle* h, void* cuda_stream);

/* Q: n3 x n3 (n3 = 3N) column-major with leading dimension ld >= n3.  Copied (and re-laid-out) into HBM. */
int  xm_set_q_dense(xm_handle* h, int n3, const double* q_colmajor, int64_t ld);
int  xm_set_q_dense_dev(xm_handle* h, int n3, const double* q_colmajor_dev, int64_t ld);
/* Block-CSR Q with bdim x bdim blocks (bdim in {3,4}); nb block rows; block values column-major per block. */
int  xm_set_q_bsr(xm_handle* h, int nb, int bdim, const int* rowptr, const int* colidx, const double* vals);

/* ---- multi-GPU: one solve partitioned by camera over `world` <= 8 GPUs of one NVSwitch node (the
 * reference is single-GPU, XM/include/Utils/memory.h:284,366).  Rank k owns a contiguous camera range (xm_partition) and
 * holds only those rows of Q; per tCG iteration the persistent kernels exchange the Q.Y operand rows and one double per
 * CTA per reduction by peer-mapped stores over NVLink — no host or NCCL call on the path.
 * Protocol (every rank): xm_create -> xm_comm_init -> exchange the 64-byte handles (any transport, e.g. a
 * torch.distributed all_gather) -> xm_comm_connect -> xm_set_q_* -> xm_trust_region* / xm_qy* / xm_op_*.  Every compute
 * call is COLLECTIVE: all ranks make the same call with the same (full-size) vector arguments; every rank receives the
 * full result.  xm_set_q_dense / xm_set_q_bsr given the whole matrix upload only the rank's rows; xm_set_q_dense_slab
 * takes the rank's row slab alone.  xm_certify on a communicator uses the iterative eigen-solver (collective). */
#define XM_IPC_HANDLE_BYTES 64
#define XM_MAX_WORLD 8
/* cameras [cam_lo, cam_hi) of `rank` when each of `world` ranks runs ctas_per_rank CTAs (pure host function, no GPU) */
int  xm_partition(int n_cameras, int world, int ctas_per_rank, int rank, int* cam_lo, int* cam_hi);
int  xm_comm_init(xm_handle* h, int rank, int world, int n_cameras, int max_r, unsigned char* ipc_handle_out /* 64 B or NULL */);
int  xm_comm_connect(xm_handle* h, const unsigned char* all_handles /* world x 64 B, rank order (one process per GPU) */);
int  xm_comm_connect_ptrs(xm_handle* h, void* const* arena_ptrs /* world pointers, rank order (one process, many GPUs) */);
int  xm_comm_disconnect(xm_handle* h);  /* unmap the peers' arenas; every rank calls it (then a host barrier) before any xm_destroy */
void* xm_comm_arena(xm_handle* h);
int  xm_comm_info(const xm_handle* h, int* rank, int* world, int* ctas_per_rank, int* cam_lo, int* cam_hi);
int  xm_comm_reset(xm_handle* h);   /* after XM_ESYNC; host-side barrier across ranks required before and after */
/* Boundary-only operand exchange of the CURRENT block-CSR operator on a communicator: n_need = remote cameras this rank unpacks per
 * exchange, n_sent = (camera, peer) pairs it pushes, n_remote = cameras owned by the other ranks (what a dense operator's full
 * all-gather moves).  xm_set_q_bsr builds the tables from the whole matrix every rank passes in. */
int  xm_comm_halo(const xm_handle* h, int* n_need, long long* n_sent, int* n_remote);
/* Reverse Cuthill-McKee camera order of a block-CSR view graph (perm_out[new] = old; pure host function, no GPU).  Ranks own
 * CONTIGUOUS camera ranges, so the order decides the cut: permute the matrix (and R, s) with it before xm_set_q_bsr and a view
 * graph with locality exchanges a few percent of the operand rows instead of all of them. */
int  xm_rcm_order(int nb, const int* rowptr, const int* colidx, int* perm_out);
/* Rows [row0, row0 + nrows) of Q only: q_slab points at element (row0, 0) of a column-major matrix with leading dimension
 * ld >= nrows (ld = n3 when it is a view into the full matrix).  Must equal the rank's range 3*cam_lo .. 3*cam_hi. */
int  xm_set_q_dense_slab(xm_handle* h, int n3, int row0, int nrows, const double* q_slab, int64_t ld);
int  xm_set_q_dense_slab_dev(xm_handle* h, int n3, int row0, int nrows, const double* q_slab_dev, int64_t ld);

/* out = alpha * Q * X ; X, out: 3N x r column-major (ld = 3N). */
int  xm_qy(xm_handle* h, int r, double alpha, const double* X, double* out);
int  xm_qy_dev(xm_handle* h, int r, double alpha, const double* X_dev, double* out_dev);
/* Measurement hook: average device time (ms, CUDA events on the handle's stream) of `iters` back-to-back launches of
 * the Q.Y kernel on the operand left in the workspace by the last xm_qy/xm_qy_dev call. */
int  xm_bench_qy(xm_handle* h, int r, int iters, double* avg_ms);   /* iters < 0: grid barrier after every product */
int  xm_bench_barrier(xm_handle* h, int r, int iters, double* avg_us);   /* iters < 0: |iters| x (barrier + operand exchange) */
int  xm_debug_counters(xm_handle* h, unsigned long long* out8);  /* raw ns counters of the last launch (profile = 1) */
int  xm_debug_trace(xm_handle* h, unsigned long long* out256);   /* (tag, ns) pairs of the last profiled solve */

/* Mirrors XMtrustregion.  R0/R_out: 3N x r col-major; s0/s_out: length N (s[0] = 1); v: length 3N or NULL when
 * ls_step == 0.  gradtol_inout is updated like the reference's by-reference gradtol.  stats/log may be NULL. */
int  xm_trust_region(xm_handle* h, int r, const double* R0, const double* s0, double lam,
                     double* gradtol_inout, double ls_step, const double* v, double max_time,
                     double* R_out, double* s_out, double* primal_out, xm_stats* stats, xm_log_rec* log);
int  xm_trust_region_dev(xm_handle* h, int r, const double* R0_dev, const double* s0_dev, double lam,
                     double* gradtol_inout, double ls_step, const double* v_dev, double max_time,
                     double* R_out_dev, double* s_out_dev, double* primal_out, xm_stats* stats, xm_log_rec* log);

/* Op-level hooks (unit-test surface; same device code as the solver phases). All R-like args 3N x r col-major. */
int  xm_op_objective(xm_handle* h, int r, const double* R, const double* s, double lam, double* f_out);
/* Riemannian gradient at (R,s): rgradR (3N x r), rgrads (N, [0]=0); also returns egrad pieces if non-NULL. */
int  xm_op_rgrad(xm_handle* h, int r, const double* R, const double* s, double lam,
                 double* rgradR, double* rgrads, double* gradnorm_out);
/* Riemannian Hessian-vector product at (R,s) along (P,ps) (ps[0] ignored -> 0): HpR (3N x r), Hps (N). */
int  xm_op_rhess(xm_handle* h, int r, const double* R, const double* s, double lam,
                 const double* P, const double* ps, double* HpR, double* Hps);
/* Retraction: Rn = MGS(R + lr*etaR) per camera ; sn = s * exp(lr*etas/s). */
int  xm_op_retract(xm_handle* h, int r, const double* R, const double* s, const double* etaR,
                   const double* etas, double lr, double* Rn, double* sn);

/* Optimality certificate (checkeig): sR = R*s is formed internally.  v_out (3N) = eigenvector of the minimum
 * eigenvalue of the dual slack.  certified_out: 1/0. */
int  xm_certify(xm_handle* h, int r, const double* R, const double* s, double lam, double primal,
                double* v_out, double* min_eig_out, double* dual_out, double* gap_out, int* certified_out);
/* Certificate with a chosen eigen-solver and a report.  XM_CERT_DENSE: cusolverDnDsyevd on the assembled 3N x 3N dual slack
 * like checkeig.h:303-318 (one GPU, dense Q, O(N^3)).  XM_CERT_ITERATIVE: block Davidson on the operator S X = Q X + blockdiag X
 * (<= 20 columns per product, block-Jacobi preconditioner, vectors resident in HBM): dense or block-CSR Q, one GPU or a
 * communicator (collective: every rank makes the same call and takes the same decision).  XM_CERT_AUTO (what xm_certify uses):
 * dense for a dense one-GPU Q with 3N <= 6000, iterative otherwise. */
enum { XM_CERT_AUTO = 0, XM_CERT_DENSE = 1, XM_CERT_ITERATIVE = 2 };
typedef struct xm_cert_info {
    int certified;          /* checkeig.h:349-368 */
    int method;             /* XM_CERT_DENSE or XM_CERT_ITERATIVE: what actually ran */
    int products;           /* Q.Y products used (each <= 20 columns) */
    int converged;          /* iterative: the r + 1 lowest Ritz pairs reached the residual tolerance */
    double min_eig, dual, gap;
    double residual;        /* iterative: largest residual norm among the r + 1 lowest Ritz pairs at exit */
    double ms;              /* device time of the whole call (CUDA events on the handle's stream) */
} xm_cert_info;
int  xm_certify_ex(xm_handle* h, int r, const double* R, const double* s, double lam, double primal, int method,
                   double* v_out, xm_cert_info* info_out);
/* 3 x 3 diagonal blocks of the operator: out[9 i + 3 a + b] = Q[3i + a, 3i + b] (host buffer, 9 N doubles). */
int  xm_op_diag_blocks(xm_handle* h, double* out9N);

/* The rank staircase of the reference's entry points in one call (XM_main.cu:180-310 solve, :312-401 solve_rank3,
 * :35-178 solve_rebuttle): rank 3 from the identity, certificate, zero-padded escalation along the escape direction, ... on
 * whatever operator the handle holds (dense / block-CSR, one GPU / a communicator — then collective).  s_init: N doubles or
 * NULL (solve_rebuttle's s_ini.bin).  R_out: capacity 3N x max(max_rank, 3) doubles, the first 3N x res->rank are the result
 * (column-major); s_out: N.  res->status: 1 certified, 2 max rank reached uncertified, 0 n/a (rank-3 mode), -2 line search failed. */
enum { XM_MODE_FULL = 0, XM_MODE_RANK3 = 1, XM_MODE_REBUTTLE = 2 };
typedef struct xm_solve_result {
    int rank, status, n_solves, certified, cert_method;
    int tcg_iters_total, qy_products_total, cert_products_total;
    double primal, gradnorm, min_eig, dual, gap;
    double solve_ms_total, cert_ms_total;
} xm_solve_result;
int  xm_solve(xm_handle* h, int mode, int max_rank, double tol, double lam, double max_time, const double* s_init,
              int cert_method, double* R_out, double* s_out, xm_solve_result* res);

/* Q assembly (create_matrix, utils/creatematrix.py:52-341) on the device: observations of the bipartite (camera, landmark)
 * graph -> the 3N x 3N SDP data matrix Q = Q1 - Vbar Lbar^-1 Vbar^T.  cam, lm: 0-based (n_obs); w: n_obs; pts: n_obs x 3 ROW-major
 * camera-frame points (host arrays).  The assembled matrix BECOMES THE HANDLE'S OPERATOR (as after xm_set_q_dense): a solve can
 * follow without a host round trip.  Q_out (host, 3N x 3N column-major) and Abar_out (host, (N + M - 1) x 3N column-major: the
 * reference's Abar.bin, dense — small problems only) may be NULL; assemble_ms_out (may be NULL): device time of the assembly. */
int  xm_create_matrix(xm_handle* h, int n_cameras, int n_landmarks, int64_t n_obs, const int* cam, const int* lm,
                      const double* w, const double* pts, double* Q_out, double* Abar_out, double* assemble_ms_out);

/* Solution recovery (recover_XM): rank-r -> 3 (top-3 eigenvectors of (sR)^T(sR)), per-camera scale ||block||_F / sqrt(3),
 * anchoring by camera 0, projection of every 3x3 block to O(3) (polar factor U V^T), global sign by majority of det, and
 * [t p] = Abar (sR)^T.  R: 3N x r col-major; s: N.  Abar: abar_rows x 3N COLUMN-MAJOR (abar_rows = N + M - 1, the
 * reference's Abar.bin) or NULL.  Outputs (host): R_out 3 x 3N col-major (camera i = columns 3i..3i+2), s_out N,
 * y_out 3 x (abar_rows + 1) col-major = [t | p] with the zero first column (required iff Abar), eig_out r eigenvalues of
 * (sR)^T(sR) descending (zeros when r == 3; may be NULL), negative_out = cameras whose block had det < 0 (may be NULL).
 * The reference's Q and lam arguments only feed a printed diagnostic and are not needed.  Works on any handle (no Q). */
int  xm_recover(xm_handle* h, int n_cameras, int r, const double* R, const double* s, const double* Abar, int64_t abar_rows,
                double* R_out, double* s_out, double* y_out, double* eig_out, int* negative_out);
/* XM^2 outlier cut: err[o] = w[o] * || p[:, lm[o]] - (s[cam[o]] R_cam[o] pt[o] + t[:, cam[o]]) ||^2 for every observation
 * (cam, lm 0-based int32; pts n_obs x 3 ROW-major; R_real 3 x 3N, t 3 x N, p 3 x M column-major as returned by xm_recover). */
int  xm_residuals(xm_handle* h, int64_t n_obs, int n_cameras, int n_landmarks, const int* cam, const int* lm, const double* pts,
                  const double* w, const double* R_real, const double* s_real, const double* t, const double* p, double* err_out);
/* v[3i..3i+2] /= s[i]  (DecentDirectionKernal) — host-side helper, trivial. */
int  xm_escape_scale(int n_cameras, double* v, const double* s);

#ifdef __cplusplus
}
#endif
#endif /* XM_B200_H */
