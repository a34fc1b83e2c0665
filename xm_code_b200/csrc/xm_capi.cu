// xm_capi.cu — host side of libxm_b200.so: the C-ABI declared in include/xm_b200.h.
// There is no CPU fallback anywhere in this file: without an sm_100 device every entry point fails with XM_ENOGPU.
#include "xm_host.h"
#include "xm_solve.cuh"
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

using namespace xm;

namespace xm {
// ------------------------------------------------------------------------------------------------ layout kernels
// Qp[i*ldq + k0 + k] = Qcm[i + ld*k]  (column-major user matrix -> padded row-major), 32x32 smem tiles.  Qcm holds `ncols`
// columns of `nrows` rows (a rank's row slab, and — on the host-upload path — one column panel starting at column k0 of the
// full matrix); columns [ncols, nwrite) of the panel are written as zero padding.
__global__ void xm_repack_q_kernel(const double* __restrict__ Qcm, long long ld, int nrows, int ncols, int k0, int nwrite,
                                   double* __restrict__ Qp, int ldq) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y * 32, bk = blockIdx.x * 32;
    for (int t = threadIdx.y; t < 32; t += blockDim.y) {        // read: consecutive threads along i (contiguous in col-major)
        const int k = bk + t, i = bi + threadIdx.x;
        tile[t][threadIdx.x] = (i < nrows && k < ncols) ? Qcm[(size_t)i + (size_t)ld * k] : 0.0;
    }
    __syncthreads();
    for (int t = threadIdx.y; t < 32; t += blockDim.y) {        // write: consecutive threads along k
        const int i = bi + t, k = bk + threadIdx.x;
        if (i < nrows && k < nwrite) Qp[(size_t)i * ldq + k0 + k] = (k < ncols) ? tile[threadIdx.x][t] : 0.0;
    }
}

// wire layout (3N x r column-major) -> camera-major operand dst[row * r + j] (block-CSR path)
__global__ void xm_operand_cam_major_kernel(const double* __restrict__ src, int n3, int r, double* __restrict__ dst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)n3 * r) {
        const int j = (int)(t / n3), row = (int)(t - (long long)j * n3);      // coalesced reads
        dst[(size_t)row * r + j] = src[t];
    }
}

}  // namespace xm

static size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

static_assert(sizeof(xm_log_rec) == sizeof(LogRec), "log record layout");
static_assert(XM_LOG_CAP == kLogCap, "log cap");
static_assert(XM_MAX_RANK == kMaxRank, "max rank");

extern "C" void xm_default_options(xm_options* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->device = 0; o->grid_ctas = 0; o->ksplit = 0; o->replicate_stale_sr = 1; o->verbose = 0;
    o->max_outer = 1000; o->max_inner = 1000; o->qy_variant = 0; o->vec_in_global = 0; o->profile = 0; o->three_barrier_tcg = 0;
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
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    {
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); h->encode_tiled = nullptr; }
    }
    if (cudaSetDevice(h->device) != cudaSuccess) { delete h; return XM_ECUDA; }
    bool ok = cudaMalloc(&h->d_stats, sizeof(DevStats)) == cudaSuccess &&
              cudaMalloc(&h->d_log, sizeof(LogRec) * kLogCap) == cudaSuccess &&
              cudaMalloc(&h->d_bar, 256) == cudaSuccess && cudaMalloc(&h->d_abort, 256) == cudaSuccess &&
              cudaMalloc(&h->d_scalar, 256) == cudaSuccess &&
              cudaMallocHost(&h->h_stats, sizeof(DevStats)) == cudaSuccess &&
              cudaMallocHost(&h->h_log, sizeof(LogRec) * kLogCap) == cudaSuccess;
    if (!ok) { xm_destroy(h); return XM_ENOMEM; }
    cudaMemset(h->d_bar, 0, 256); cudaMemset(h->d_abort, 0, 256);
    *out = h;
    return XM_OK;
}

extern "C" int xm_destroy(xm_handle* h) {
    if (!h) return XM_OK;
    cudaSetDevice(h->device);
    cudaFree(h->Qp); cudaFree(h->Qstage); cudaFree(h->bsr_rowptr); cudaFree(h->bsr_col); cudaFree(h->bsr_val); cudaFree(h->d_peer_mask); cudaFree(h->d_need_cams);
    cudaFree(h->ws); cudaFree(h->d_stats); cudaFree(h->d_log); cudaFree(h->d_bar); cudaFree(h->d_abort); cudaFree(h->d_scalar);
    cudaFree(h->io_R0); cudaFree(h->io_s0); cudaFree(h->io_v); cudaFree(h->io_Rout); cudaFree(h->io_sout); cudaFree(h->io_P); cudaFree(h->io_ps);
    if (h->cublas) cublasDestroy((cublasHandle_t)h->cublas);
    if (h->cusolver) cusolverDnDestroy((cusolverDnHandle_t)h->cusolver);
    cudaFreeHost(h->h_stats); cudaFreeHost(h->h_log);
    for (int w = 0; w < kMaxWorld; ++w) if (h->peer_ipc[w] && h->peer_arena[w]) cudaIpcCloseMemHandle(h->peer_arena[w]);
    cudaFree(h->arena);
    cudaGetLastError();
    delete h;
    return XM_OK;
}

extern "C" int xm_set_stream(xm_handle* h, void* s) {
    if (!h) return XM_EINVAL;
    h->stream = (cudaStream_t)s;
    if (h->cublas) cublasSetStream((cublasHandle_t)h->cublas, h->stream);
    if (h->cusolver) cusolverDnSetStream((cusolverDnHandle_t)h->cusolver, h->stream);
    return XM_OK;
}

// the handle's cuBLAS / cuSOLVER handles (created on first use, bound to the handle's stream, destroyed with the handle)
int xm_internal_libs(xm_handle* h, void** cublas_out, void** cusolver_out) {
    if (!h->cublas) {
        cublasHandle_t cb = nullptr;
        if (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS) { h->err = "cublasCreate failed"; return XM_ECUDA; }
        h->cublas = cb;
    }
    if (!h->cusolver) {
        cusolverDnHandle_t cs = nullptr;
        if (cusolverDnCreate(&cs) != CUSOLVER_STATUS_SUCCESS) { h->err = "cusolverDnCreate failed"; return XM_ECUDA; }
        h->cusolver = cs;
    }
    cublasSetStream((cublasHandle_t)h->cublas, h->stream);
    cusolverDnSetStream((cusolverDnHandle_t)h->cusolver, h->stream);
    if (cublas_out) *cublas_out = h->cublas;
    if (cusolver_out) *cusolver_out = h->cusolver;
    return XM_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU communicator
// Camera partition of SURVEY.md §8e.  The job runs on world * G CTAs; global CTA g owns the cameras [g N / GT, (g+1) N / GT),
// so rank k (CTAs [kG, (k+1)G)) owns the contiguous range below and holds exactly those rows of Q.
extern "C" int xm_partition(int n_cameras, int world, int ctas_per_rank, int rank, int* cam_lo, int* cam_hi) {
    if (n_cameras <= 0 || world < 1 || world > kMaxWorld || ctas_per_rank < 1 || rank < 0 || rank >= world) return XM_EINVAL;
    const long long GT = (long long)world * ctas_per_rank;
    if (cam_lo) *cam_lo = (int)(((long long)rank * ctas_per_rank * n_cameras) / GT);
    if (cam_hi) *cam_hi = (int)(((long long)(rank + 1) * ctas_per_rank * n_cameras) / GT);
    return XM_OK;
}

extern "C" int xm_comm_init(xm_handle* h, int rank, int world, int n_cameras, int max_r, unsigned char* ipc_handle_out) {
    if (!h) return XM_EINVAL;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n_cameras < world || max_r < 3 || max_r > XM_MAX_RANK) {
        h->err = "xm_comm_init: bad rank/world/camera count/max rank"; return XM_EINVAL;
    }
    if (h->arena) { h->err = "communicator already initialised on this handle"; return XM_EINVAL; }
    XM_CUDA(h, cudaSetDevice(h->device));
    int G = h->opt.grid_ctas > 0 ? h->opt.grid_ctas : h->num_sm;
    G = std::max(1, std::min(std::min(G, h->num_sm), n_cameras / world));
    const size_t n3 = 3 * (size_t)n_cameras, ldq = (n3 + 63) / 64 * 64;
    size_t off = 0;
    h->off_bar = off; off += 256;
    h->off_abort = off; off += 256;
    h->off_partials = off; off += align_up_sz((size_t)kPartialBufs * (G + 1) * kPartialStride * sizeof(double), 256);
    h->off_ll = off; off += align_up_sz((size_t)kPartialBufs * 4 * kMaxWorld * sizeof(unsigned long long), 256);
    h->off_slots = off; off += align_up_sz((size_t)kPartialBufs * G * sizeof(ulonglong2), 256);
    h->off_xt = off; off += align_up_sz((size_t)max_r * ldq * sizeof(double), 256);
    h->off_xtll = off; off += align_up_sz((size_t)max_r * ldq * sizeof(ulonglong2), 256);
    h->off_outR = off; off += align_up_sz(n3 * max_r * sizeof(double), 256);
    h->off_outS = off; off += align_up_sz((size_t)n_cameras * sizeof(double), 256);
    if (cudaMalloc(&h->arena, off) != cudaSuccess) { cudaGetLastError(); h->arena = nullptr; h->err = "communicator arena cudaMalloc failed"; return XM_ENOMEM; }
    h->arena_bytes = off;
    XM_CUDA(h, cudaMemset(h->arena, 0, off));
    XM_CUDA(h, cudaMemset(h->d_bar, 0, 256));
    XM_CUDA(h, cudaDeviceSynchronize());
    if (ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == XM_IPC_HANDLE_BYTES, "ipc handle size");
        cudaIpcMemHandle_t mh;
        XM_CUDA(h, cudaIpcGetMemHandle(&mh, h->arena));
        memcpy(ipc_handle_out, &mh, sizeof(mh));
    }
    h->world = world; h->rank = rank; h->comm_G = G; h->comm_N = n_cameras; h->comm_maxr = max_r;
    h->comm_connected = (world == 1); h->comm_broken = false;
    for (int w = 0; w < kMaxWorld; ++w) { h->peer_arena[w] = nullptr; h->peer_ipc[w] = false; }
    h->peer_arena[rank] = h->arena;
    xm_partition(n_cameras, world, G, rank, &h->cam0, &h->cam1);
    // an operator uploaded before the communicator existed covers the wrong rows
    h->N = 0; h->n3 = 0;
    return XM_OK;
}

// one process per GPU: `all_handles` = the world x 64-byte handles returned by xm_comm_init on every rank, in rank order
extern "C" int xm_comm_connect(xm_handle* h, const unsigned char* all_handles) {
    if (!h || !all_handles || !h->arena) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    for (int w = 0; w < h->world; ++w) {
        if (w == h->rank) continue;
        cudaIpcMemHandle_t mh;
        memcpy(&mh, all_handles + (size_t)w * XM_IPC_HANDLE_BYTES, sizeof(mh));
        void* p = nullptr;
        XM_CUDA(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->peer_arena[w] = (char*)p; h->peer_ipc[w] = true;
    }
    h->comm_connected = true;
    return XM_OK;
}

// one process driving several GPUs: `arena_ptrs` = xm_comm_arena() of every rank's handle, in rank order
extern "C" int xm_comm_connect_ptrs(xm_handle* h, void* const* arena_ptrs) {
    if (!h || !arena_ptrs || !h->arena) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    for (int w = 0; w < h->world; ++w) {
        if (w == h->rank) continue;
        cudaPointerAttributes at;
        XM_CUDA(h, cudaPointerGetAttributes(&at, arena_ptrs[w]));
        if (at.device == h->device) { h->peer_arena[w] = (char*)arena_ptrs[w]; continue; }   // loop-back member on the same GPU
        int can = 0;
        XM_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, at.device));
        if (!can) { h->err = "no peer access between the communicator's devices"; return XM_EUNSUPPORTED; }
        cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return XM_ECUDA; }
        cudaGetLastError();
        h->peer_arena[w] = (char*)arena_ptrs[w];
    }
    h->comm_connected = true;
    return XM_OK;
}

// Unmap the peers' arenas (before any rank frees its own: the importer must close first).  Collective by convention:
// every rank disconnects, the ranks meet at a host barrier, then handles may be destroyed.
extern "C" int xm_comm_disconnect(xm_handle* h) {
    if (!h) return XM_EINVAL;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int w = 0; w < kMaxWorld; ++w) {
        if (h->peer_ipc[w] && h->peer_arena[w]) cudaIpcCloseMemHandle(h->peer_arena[w]);
        if (w != h->rank) h->peer_arena[w] = nullptr;
        h->peer_ipc[w] = false;
    }
    cudaGetLastError();
    if (h->world > 1) h->comm_connected = false;
    return XM_OK;
}

extern "C" void* xm_comm_arena(xm_handle* h) { return h ? (void*)h->arena : nullptr; }

extern "C" int xm_comm_info(const xm_handle* h, int* rank, int* world, int* ctas_per_rank, int* cam_lo, int* cam_hi) {
    if (!h) return XM_EINVAL;
    if (rank) *rank = h->rank;
    if (world) *world = h->world;
    if (ctas_per_rank) *ctas_per_rank = h->comm_G;
    if (cam_lo) *cam_lo = h->cam0;
    if (cam_hi) *cam_hi = h->cam1;
    return XM_OK;
}

// After XM_ESYNC the ranks' barrier epochs may differ.  Call on EVERY rank, with a host-side barrier across the ranks
// before (no kernel in flight anywhere) and after (nobody launches into a counter that is about to be zeroed).
extern "C" int xm_comm_reset(xm_handle* h) {
    if (!h || !h->arena) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    XM_CUDA(h, cudaDeviceSynchronize());
    XM_CUDA(h, cudaMemset(h->arena, 0, h->arena_bytes));     // counters, abort flag, every tagged word
    XM_CUDA(h, cudaMemset(h->d_bar, 0, 256));
    XM_CUDA(h, cudaDeviceSynchronize());
    h->comm_broken = false;
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

// rows of the operator this handle keeps: everything, or the slab of its cameras once a communicator is attached
static void owned_cameras(const xm_handle* h, int N, int* c0, int* c1) {
    if (h->world > 1) { *c0 = h->cam0; *c1 = h->cam1; } else { *c0 = 0; *c1 = N; }
}

// q_slab: first element of row `row0` of the column-major n3-column matrix (leading dimension ld), nrows rows
static int set_q_common(xm_handle* h, int n3, int row0, int nrows, const double* q_slab, int64_t ld, bool from_device) {
    XmRange nvtx_range("xm_set_q_dense");
    if (!h || !q_slab || n3 <= 0 || n3 % 3 != 0 || ld < nrows || nrows <= 0) return XM_EINVAL;
    if (h->world > 1 && h->comm_N * 3 != n3) { h->err = "Q size differs from the communicator's camera count"; return XM_EINVAL; }
    int c0, c1;
    owned_cameras(h, n3 / 3, &c0, &c1);
    if (row0 != 3 * c0 || nrows != 3 * (c1 - c0)) { h->err = "row slab does not match this rank's camera range"; return XM_EINVAL; }
    XM_CUDA(h, cudaSetDevice(h->device));
    const int ldq = (n3 + 63) / 64 * 64;
    int rc = ensure(h, &h->Qp, &h->Qp_cap, (size_t)nrows * ldq * sizeof(double));
    if (rc) return rc;
    if (from_device) {
        dim3 blk(32, 8), grd((ldq + 31) / 32, (nrows + 31) / 32);
        xm_repack_q_kernel<<<grd, blk, 0, h->stream>>>(q_slab, (long long)ld, nrows, n3, 0, ldq, h->Qp, ldq);
        h->launches++;
        XM_CUDA(h, cudaGetLastError());
    } else {
        // host matrix: column panels of at most ~128 MB go through ONE bounded staging buffer (H2D copy, then the re-layout
        // kernel writes the panel's columns of the padded row-major copy) — no second full-size copy of Q in HBM
        const size_t col_bytes = (size_t)nrows * sizeof(double);
        int kc = (int)std::max<size_t>(32, ((size_t)128 << 20) / col_bytes / 32 * 32);
        kc = std::min(kc, (n3 + 31) / 32 * 32);
        rc = ensure(h, &h->Qstage, &h->Qstage_cap, (size_t)kc * col_bytes);
        if (rc) return rc;
        for (int k0 = 0; k0 < ldq; k0 += kc) {
            const int ncopy = std::max(0, std::min(kc, n3 - k0));            // real columns in this panel (the rest is padding)
            const int nwrite = std::min(kc, ldq - k0);
            if (ncopy > 0)
                XM_CUDA(h, cudaMemcpy2DAsync(h->Qstage, col_bytes, q_slab + (size_t)ld * k0, (size_t)ld * sizeof(double), col_bytes, ncopy,
                                             cudaMemcpyHostToDevice, h->stream));
            dim3 blk(32, 8), grd((nwrite + 31) / 32, (nrows + 31) / 32);
            xm_repack_q_kernel<<<grd, blk, 0, h->stream>>>(h->Qstage, (long long)nrows, nrows, ncopy, k0, nwrite, h->Qp, ldq);
            XM_CUDA(h, cudaGetLastError());
            h->launches++;
        }
        XM_CUDA(h, cudaStreamSynchronize(h->stream));      // "copied": the caller may reuse its buffer when this returns
    }
    h->n3 = n3; h->N = n3 / 3; h->ldq = ldq; h->is_bsr = false;
    h->cam0 = c0; h->cam1 = c1;
    return XM_OK;
}
// full matrix given: a rank of a communicator uploads only the rows of its own cameras
static int set_q_full(xm_handle* h, int n3, const double* q, int64_t ld, bool from_device) {
    if (!h || !q || n3 <= 0 || n3 % 3 != 0 || ld < n3) return XM_EINVAL;
    int c0, c1;
    owned_cameras(h, n3 / 3, &c0, &c1);
    return set_q_common(h, n3, 3 * c0, 3 * (c1 - c0), q + 3 * c0, ld, from_device);
}
extern "C" int xm_set_q_dense(xm_handle* h, int n3, const double* q, int64_t ld) { return set_q_full(h, n3, q, ld, false); }
extern "C" int xm_set_q_dense_dev(xm_handle* h, int n3, const double* q, int64_t ld) { return set_q_full(h, n3, q, ld, true); }
extern "C" int xm_set_q_dense_slab(xm_handle* h, int n3, int row0, int nrows, const double* q_slab, int64_t ld) {
    return set_q_common(h, n3, row0, nrows, q_slab, ld, false);
}
extern "C" int xm_set_q_dense_slab_dev(xm_handle* h, int n3, int row0, int nrows, const double* q_slab, int64_t ld) {
    return set_q_common(h, n3, row0, nrows, q_slab, ld, true);
}

extern "C" int xm_set_q_bsr(xm_handle* h, int nb, int bdim, const int* rowptr, const int* colidx, const double* vals) {
    if (!h || !rowptr || !colidx || !vals || nb <= 0 || (bdim != 3 && bdim != 4)) return XM_EINVAL;
    if (h->world > 1 && h->comm_N != nb) { h->err = "Q size differs from the communicator's camera count"; return XM_EINVAL; }
    XM_CUDA(h, cudaSetDevice(h->device));
    // a rank of a communicator keeps only the block rows of its own cameras (the caller passes the whole matrix)
    int c0, c1;
    owned_cameras(h, nb, &c0, &c1);
    const long long bb0 = rowptr[c0], nnzb = (long long)rowptr[c1] - bb0;
    std::vector<int> rp((size_t)(c1 - c0) + 1);
    for (int i = c0; i <= c1; ++i) rp[i - c0] = (int)(rowptr[i] - bb0);
    // re-block on the host into 4x4 row-major padded blocks (128 B each); for bdim==4 only the leading 3x3 acts on
    // the rotation rows (the 4th row/col belongs to translations, which the BM path has already eliminated).
    std::vector<double> blk((size_t)nnzb * 16, 0.0);
    for (long long b = 0; b < nnzb; ++b)
        for (int cc = 0; cc < bdim; ++cc)
            for (int rr = 0; rr < bdim; ++rr)
                blk[(size_t)b * 16 + rr * 4 + cc] = vals[(size_t)(bb0 + b) * bdim * bdim + (size_t)cc * bdim + rr];
    cudaFree(h->bsr_rowptr); cudaFree(h->bsr_col); cudaFree(h->bsr_val);
    h->bsr_rowptr = nullptr; h->bsr_col = nullptr; h->bsr_val = nullptr;
    XM_CUDA(h, cudaMalloc(&h->bsr_rowptr, sizeof(int) * rp.size()));
    XM_CUDA(h, cudaMalloc(&h->bsr_col, sizeof(int) * std::max(nnzb, 1LL)));
    XM_CUDA(h, cudaMalloc(&h->bsr_val, sizeof(double) * 16 * std::max(nnzb, 1LL)));
    XM_CUDA(h, cudaMemcpy(h->bsr_rowptr, rp.data(), sizeof(int) * rp.size(), cudaMemcpyHostToDevice));
    XM_CUDA(h, cudaMemcpy(h->bsr_col, colidx + bb0, sizeof(int) * nnzb, cudaMemcpyHostToDevice));
    XM_CUDA(h, cudaMemcpy(h->bsr_val, blk.data(), sizeof(double) * 16 * nnzb, cudaMemcpyHostToDevice));
    h->bsr_bdim = bdim; h->is_bsr = true;
    h->N = nb; h->n3 = 3 * nb; h->ldq = (h->n3 + 63) / 64 * 64;
    h->cam0 = c0; h->cam1 = c1;
    cudaFree(h->d_peer_mask); cudaFree(h->d_need_cams);
    h->d_peer_mask = nullptr; h->d_need_cams = nullptr; h->n_need = 0; h->halo_sent = 0;
    if (h->world > 1) {
        // boundary-only exchange: which ranks reference which cameras (the caller passed the whole matrix, so every rank can
        // work out the same table); this rank's own list of remote cameras to unpack
        std::vector<unsigned char> mask((size_t)nb, 0);
        std::vector<int> need;
        for (int w = 0; w < h->world; ++w) {
            int lo, hi;
            xm_partition(nb, h->world, h->comm_G, w, &lo, &hi);
            for (long long b = rowptr[lo]; b < rowptr[hi]; ++b) {
                const int c = colidx[b];
                if (c < 0 || c >= nb) { h->err = "block column index out of range"; return XM_EINVAL; }
                mask[c] |= (unsigned char)(1u << w);
            }
        }
        for (int c = 0; c < nb; ++c) {
            if ((c < c0 || c >= c1) && ((mask[c] >> h->rank) & 1u)) need.push_back(c);
            if (c >= c0 && c < c1) for (int w = 0; w < h->world; ++w) if (w != h->rank && ((mask[c] >> w) & 1u)) h->halo_sent++;
        }
        h->n_need = (int)need.size();
        XM_CUDA(h, cudaMalloc(&h->d_peer_mask, (size_t)nb));
        XM_CUDA(h, cudaMalloc(&h->d_need_cams, sizeof(int) * std::max<size_t>(need.size(), 1)));
        XM_CUDA(h, cudaMemcpy(h->d_peer_mask, mask.data(), (size_t)nb, cudaMemcpyHostToDevice));
        if (!need.empty()) XM_CUDA(h, cudaMemcpy(h->d_need_cams, need.data(), sizeof(int) * need.size(), cudaMemcpyHostToDevice));
    }
    return XM_OK;
}

// boundary-only exchange of the current block-CSR operator: remote cameras this rank unpacks per exchange, (camera, peer) pairs it
// pushes, and the remote cameras there are (what the full all-gather of a dense operator moves)
extern "C" int xm_comm_halo(const xm_handle* h, int* n_need, long long* n_sent, int* n_remote) {
    if (!h) return XM_EINVAL;
    const bool on = h->world > 1 && h->is_bsr && h->d_peer_mask;
    if (n_need) *n_need = on ? h->n_need : (h->world > 1 ? h->N - (h->cam1 - h->cam0) : 0);
    if (n_sent) *n_sent = on ? h->halo_sent : (long long)(h->cam1 - h->cam0) * (h->world - 1);
    if (n_remote) *n_remote = h->world > 1 ? h->N - (h->cam1 - h->cam0) : 0;
    return XM_OK;
}

// Reverse Cuthill-McKee order of the cameras of a block-CSR view graph ("simple BFS/RCM bands"): perm_out[new] = old.
// The library partitions CONTIGUOUS camera ranges over the ranks, so the camera order decides the cut: after this ordering a view
// graph with locality becomes banded and the boundary-only exchange moves a few percent of the rows.  Pure host function.
extern "C" int xm_rcm_order(int nb, const int* rowptr, const int* colidx, int* perm_out) {
    if (nb <= 0 || !rowptr || !colidx || !perm_out) return XM_EINVAL;
    std::vector<int> deg(nb), order; std::vector<char> seen(nb, 0);
    for (int i = 0; i < nb; ++i) deg[i] = rowptr[i + 1] - rowptr[i];
    std::vector<int> by_deg(nb);
    for (int i = 0; i < nb; ++i) by_deg[i] = i;
    std::stable_sort(by_deg.begin(), by_deg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
    order.reserve(nb);
    std::vector<int> nbrs;
    for (int s = 0; s < nb; ++s) {                       // one BFS per connected component, started at a minimum-degree camera
        const int root = by_deg[s];
        if (seen[root]) continue;
        seen[root] = 1;
        size_t head = order.size();
        order.push_back(root);
        while (head < order.size()) {
            const int u = order[head++];
            nbrs.clear();
            for (int b = rowptr[u]; b < rowptr[u + 1]; ++b) { const int v = colidx[b]; if (v >= 0 && v < nb && !seen[v]) { seen[v] = 1; nbrs.push_back(v); } }
            std::stable_sort(nbrs.begin(), nbrs.end(), [&](int a, int b) { return deg[a] < deg[b]; });
            order.insert(order.end(), nbrs.begin(), nbrs.end());
        }
    }
    for (int k = 0; k < nb; ++k) perm_out[k] = order[nb - 1 - k];
    return XM_OK;
}

// ------------------------------------------------------------------------------------------------ launch planning
struct Plan { int RP, NT, NW, W, cpw, NSW, G, GT, KS, CB; int use_tma, KC, ST, nbmax, nchunks, stage_doubles; int box_nb[3]; size_t dyn_smem; int vec_smem, cpc; size_t vec_bytes; int nprod, NWC; };

static int rank_pad(int r) {
    static const int pads[] = {3, 4, 5, 6, 8, 10, 12, 16, 20};
    for (int p : pads) if (r <= p) return p;
    return -1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D FP64 tensor map over a row-major [rows x cols] array with row pitch `pitch` elements; box = [box_rows x box_cols]
static int make_map(xm_handle* h, CUtensorMap* m, const double* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_cols, uint32_t box_rows) {
    if (!h->encode_tiled) { h->err = "cuTensorMapEncodeTiled unavailable"; return XM_ECUDA; }
    cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {pitch * sizeof(double)};
    cuuint32_t box[2] = {box_cols, box_rows}; cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return XM_ECUDA; }
    return XM_OK;
}

static Plan make_plan(const xm_handle* h, int r, int allow_tma = 1) {
    Plan p{};
    p.RP = rank_pad(r);
    p.NT = (p.RP <= 6) ? 512 : 256;                    // dense sweeps hold 3 x RP accumulators per lane (6 x RP at RP = 8, 10: two cameras per warp)
    if (p.RP >= 8 && p.RP <= 10) { if (const char* e = getenv("XM_TUNE_DENSE_NT")) { if (atoi(e) == 512) p.NT = 512; } }      // A/B hook: one camera per warp
    if (h->is_bsr) {
        // block-CSR holds 3 accumulators per lane whatever the rank; the operand gather is latency-bound, so resident warps
        // matter more than registers: 1024 threads (32 warps, <= 64 registers, 16-block chunks, 2 gathers in flight per sub-warp)
        // measured 8 / 21 / 23 % faster than 512 threads at r = 5 / 10 / 20 in the standalone experiment (profiles/r02_bsr_tune.txt)
        p.NT = (r <= 5) ? 512 : 1024;      // in-library (profiles/r02_bsr_er100k_*): 512 threads win at r <= 5, 1024 at r >= 10
        if (const char* e = getenv("XM_TUNE_BSR_NT")) { const int v = atoi(e); if (v == 512 || v == 1024) p.NT = v; }             // A/B hook
    }
    p.NW = p.NT / 32;
    p.W = 4; while (p.W < r) p.W <<= 1;
    p.cpw = 32 / p.W;
    p.NSW = p.NW * p.cpw;
    int G = h->opt.grid_ctas > 0 ? h->opt.grid_ctas : h->num_sm;
    G = std::min(G, h->num_sm);                          // cooperative launch: one CTA per SM
    G = std::max(1, std::min(G, h->N));
    if (h->world > 1) G = h->comm_G;                     // fixed when the communicator was created (identical on every rank)
    p.G = G;
    const int GT = G * h->world;                         // CTAs of the whole job: cameras are dealt out over all of them
    const int cpc = (h->N + GT - 1) / GT;
    p.use_tma = (allow_tma && !h->is_bsr && h->opt.qy_variant != 1 && h->encode_tiled) ? 1 : 0;
    p.nprod = p.use_tma ? 1 : 0;                        // TMA path: the last nprod warps are producers (1 suffices)
    if (const char* e = getenv("XM_TUNE_NPROD")) { int v = atoi(e); if (p.use_tma && v >= 1 && v <= 4) p.nprod = v; }     // tuning hook
    const int nwork = p.NW - p.nprod;                   // warps that stream
    p.NWC = nwork;
    const int cams = p.use_tma ? dense_cams_per_warp(p.RP, p.NT) : 1;      // cameras per consumer warp (xm_device.cuh)
    p.KC = kKC;     // compile-time chunk width of the TMA ring (xm_device.cuh)
    p.nchunks = (h->ldq + p.KC - 1) / p.KC;
    // k-split: the largest divisor KS of nwork with KS * cpc <= nwork (or the caller's cap), at most one chunk/step each
    const int kmax = p.use_tma ? p.nchunks : (h->is_bsr ? nwork : std::max(1, h->ldq / 64));
    int KS = 1;
    for (int k = 1; k <= nwork; ++k)
        if (nwork % k == 0 && k <= kmax && ((h->opt.ksplit > 0) ? (k <= h->opt.ksplit) : (k * cpc <= nwork * cams))) KS = k;
    p.KS = KS; p.CB = nwork / KS * cams;
    p.cpc = cpc; p.GT = GT;
    // per-CTA state vectors in shared memory when they are small (kills the L2 round trips of every per-camera phase)
    size_t budget = (size_t)std::min(h->smem_optin, 227 * 1024) - 12 * 1024;               // static smem + slack
    p.vec_bytes = (((size_t)(kNumVecR * 3 * r + kNumVecS + 6) * cpc * sizeof(double)) + 127) / 128 * 128;
    p.vec_smem = (h->opt.vec_in_global == 0 && p.vec_bytes <= 64 * 1024) ? 1 : 0;
    if (p.vec_smem) budget -= p.vec_bytes;
    p.dyn_smem = (p.vec_smem ? p.vec_bytes : 0) + 256;
    if (h->is_bsr) p.dyn_smem += (size_t)p.NW * 2 * ((p.NT == 1024 ? 16 : 32) * 128 + 8);      // per-warp staging of the block chunks (xm_device.cuh: bsr_issue)
    if (p.use_tma) {
        // batch heights that occur: CTAs own q or q+1 cameras, swept in EVEN batches (xm_device.cuh: Batches) — at most three distinct sizes
        const int q = h->N / GT;
        int nsz = 0;
        auto add = [&](int v) { if (v <= 0) return; for (int t = 0; t < nsz; ++t) if (p.box_nb[t] == v) return; if (nsz < 3) p.box_nb[nsz++] = v; };
        for (int n = q; n <= q + 1; ++n) {
            if (n <= 0) continue;
            const int nbat = (n + p.CB - 1) / p.CB, base = n / nbat;
            add(base + (n % nbat ? 1 : 0)); add(base);
        }
        while (nsz < 3) { p.box_nb[nsz] = p.box_nb[0]; ++nsz; }
        p.nbmax = std::max(p.box_nb[0], std::max(p.box_nb[1], p.box_nb[2]));
        p.stage_doubles = (3 * p.nbmax + p.RP) * p.KC;   // operand area sized for the padded rank (consumers read RP rows)
        int ST = (int)((budget - 1024) / ((size_t)p.stage_doubles * sizeof(double)));
        ST = std::min(ST, 24);
        if (const char* e = getenv("XM_TUNE_ST")) { int v = atoi(e); if (v >= 2) ST = std::min(ST, v); }                      // tuning hook
        if (ST < 2) return make_plan(h, r, 0);          // ring does not fit: direct streaming loads
        p.ST = ST;
        p.dyn_smem += (size_t)ST * p.stage_doubles * sizeof(double) + 3 * ST * sizeof(unsigned long long);
    }
    return p;
}

static size_t align_up(size_t x, size_t a) { return align_up_sz(x, a); }

// carve the workspace for (N, r, G); zero it when the geometry changed (operand pad rows must be zero)
static int carve(xm_handle* h, int r, const Plan& p) {
    const size_t N = h->N, n3 = h->n3, ldq = h->ldq;
    const size_t vecR = align_up(n3 * r * sizeof(double), 256), vecS = align_up(N * sizeof(double), 256);
    const size_t total = kNumVecR * vecR + align_up(N * 6 * sizeof(double), 256) + kNumVecS * vecS + align_up((size_t)r * ldq * sizeof(double), 256) +
                         align_up((size_t)kPartialBufs * (p.G + 1) * kPartialStride * sizeof(double), 256);
    bool fresh = false;
    if (h->ws_cap < total) {
        // a member of a communicator sizes the workspace for the communicator's maximum rank at once: growing it later would
        // cudaFree (a device-wide synchronisation) while a peer's persistent kernel may already be waiting for this rank
        const size_t rcap = (h->world > 1) ? (size_t)std::max(r, h->comm_maxr) : (size_t)r;
        const size_t want = std::max(total, kNumVecR * align_up(n3 * rcap * sizeof(double), 256) + align_up(N * 6 * sizeof(double), 256) + kNumVecS * vecS +
                                                align_up(rcap * ldq * sizeof(double), 256) +
                                                align_up((size_t)kPartialBufs * (p.G + 1) * kPartialStride * sizeof(double), 256));
        if (h->ws) cudaFree(h->ws);
        h->ws = nullptr; h->ws_cap = 0;
        if (cudaMalloc(&h->ws, want) != cudaSuccess) { cudaGetLastError(); h->err = "workspace cudaMalloc failed"; return XM_ENOMEM; }
        h->ws_cap = want; fresh = true;
    }
    if (fresh || h->ws_r != r || h->ws_N != (int)N || h->ws_G != p.GT || h->ws_ldq != (int)ldq || h->ws_bsr != (int)h->is_bsr) {
        XM_CUDA(h, cudaMemsetAsync(h->ws, 0, total, h->stream));
        h->ws_r = r; h->ws_N = (int)N; h->ws_G = p.GT; h->ws_ldq = (int)ldq; h->ws_bsr = (int)h->is_bsr;
    }
    char* q = h->ws;
    Dev& d = h->dev;
    d = Dev{};
    d.rbase = (double*)q; d.rstride = (long long)(vecR / sizeof(double)); q += kNumVecR * vecR;
    d.S6 = (double*)q; q += align_up(N * 6 * sizeof(double), 256);
    d.sbase = (double*)q; d.sstride = (long long)(vecS / sizeof(double)); q += kNumVecS * vecS;
    d.Xt = (double*)q; q += align_up((size_t)r * ldq * sizeof(double), 256);
    d.partials = (double*)q;
    d.N = (int)N; d.r = r; d.n3 = (int)n3; d.ldq = (int)ldq;
    d.x_cam_major = h->is_bsr ? 1 : 0;
    d.e_rec = h->opt.three_barrier_tcg ? 0 : 1;           // two-barrier tCG iteration (E recurrence) unless the caller asks for the reference's order
    if (const char* e = getenv("XM_TUNE_EREC")) d.e_rec = atoi(e) ? 1 : 0;                        // A/B hook
    d.bsr_stage = 0;                        // measured best on B200 (profiles/r01_bsr_qy.md): bulk-TMA chunks
    d.bsr_chunk = (p.NT == 1024) ? 16 : 32; d.bsr_k = (p.NT == 1024) ? 2 : 4;
    if (const char* e = getenv("XM_TUNE_BSR")) { const int v = atoi(e); d.bsr_stage = v & 1; if ((v >> 1) & 1) d.bsr_k = 8; }      // A/B hook
    if (const char* e = getenv("XM_TUNE_BSR_K")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) d.bsr_k = v; }
    d.Q = h->is_bsr ? nullptr : h->Qp;
    if (h->is_bsr) { d.bsr_rowptr = h->bsr_rowptr; d.bsr_col = h->bsr_col; d.bsr_val = h->bsr_val; d.bsr_bdim = h->bsr_bdim; }   // else null: dense
    d.G = p.G; d.NW = p.NW; d.KS = p.KS; d.CB = p.CB; d.W = p.W; d.cpw = p.cpw; d.NSW = p.NSW;
    d.rank = h->rank; d.world = h->world; d.GT = p.GT; d.g0 = h->rank * p.G; d.cam0 = h->cam0; d.row0 = 3 * h->cam0;
    d.bar = h->d_bar; d.abort_flag = h->d_abort; d.epoch_store = h->d_bar + 16;
    if (h->world > 1) {       // exchange buffers live in the peer-mapped arenas (same layout on every rank)
        d.Xt = (double*)(h->arena + h->off_xt); d.partials = (double*)(h->arena + h->off_partials);
        d.bar = (unsigned long long*)(h->arena + h->off_bar); d.abort_flag = (int*)(h->arena + h->off_abort);
        d.ll = (unsigned long long*)(h->arena + h->off_ll);
        d.slots_ll = (ulonglong2*)(h->arena + h->off_slots); d.XtLL = (ulonglong2*)(h->arena + h->off_xtll);
        d.nown = 3 * (h->cam1 - h->cam0);
        // protocol switch (like NCCL's LL vs Simple): tagged words double the bytes but need no fence; above ~2 MB per rank and
        // exchange the NVLink time of the extra bytes outweighs the 3.5 us fence
        const size_t pushed_rows = (h->is_bsr && h->d_peer_mask) ? (size_t)3 * h->halo_sent : (size_t)d.nown * (h->world - 1);
        d.push_plain = (pushed_rows * r * sizeof(double) > (size_t)(2 << 20)) ? 1 : 0;
        if (const char* e = getenv("XM_TUNE_PUSH")) d.push_plain = atoi(e) ? 1 : 0;                 // A/B hook
        for (int w = 0; w < h->world; ++w) {
            d.Xt_peer[w] = (double*)(h->peer_arena[w] + h->off_xt); d.ll_peer[w] = (unsigned long long*)(h->peer_arena[w] + h->off_ll);
            d.XtLL_peer[w] = (ulonglong2*)(h->peer_arena[w] + h->off_xtll);
            d.abort_peer[w] = (int*)(h->peer_arena[w] + h->off_abort);
        }
    } else {
        d.Xt_peer[0] = d.Xt; d.abort_peer[0] = d.abort_flag;
    }
    if (h->world > 1 && h->is_bsr && h->d_peer_mask && !getenv("XM_TUNE_FULL_EXCHANGE")) { d.peer_mask = h->d_peer_mask; d.need_cams = h->d_need_cams; d.n_need = h->n_need; }
    {
        double scale = 1.0;       // compute-sanitizer / debugger runs are 10-100x slower: XM_WATCHDOG_SCALE stretches the abort deadline
        if (const char* e = getenv("XM_WATCHDOG_SCALE")) scale = std::max(1.0, atof(e));
        d.watchdog_ns = (unsigned long long)((h->world > 1 ? 30e9 : 4e9) * scale);    // multi-GPU: generous — the ranks' hosts launch independently
    }
    d.vec_smem = p.vec_smem; d.cpc_max = p.cpc; d.profile = h->opt.profile;
    d.nprod = p.nprod; d.NWC = p.NWC;
    d.l2_prefetch = 0;     // measured: L2 prefetch shortens the Q.Y phase but lengthens the barriers by as much (profiles/r01_sweep_l2_prefetch.txt)
    if (const char* e = getenv("XM_TUNE_PF")) d.l2_prefetch = std::max(0, atoi(e));          // tuning hook
    d.use_tma = p.use_tma; d.KC = p.KC; d.ST = p.ST; d.nbmax = p.nbmax; d.nchunks = p.nchunks; d.stage_doubles = p.stage_doubles;
    if (p.use_tma) {
        int rc = make_map(h, &h->mapX, d.Xt, (uint64_t)ldq, (uint64_t)r, (uint64_t)ldq, (uint32_t)p.KC, (uint32_t)r);
        if (rc) return rc;
        for (int t = 0; t < 3; ++t) {
            rc = make_map(h, &h->mapQ[t], h->Qp, (uint64_t)ldq, (uint64_t)(3 * (h->cam1 - h->cam0)), (uint64_t)ldq, (uint32_t)p.KC, (uint32_t)(3 * p.box_nb[t]));
            if (rc) return rc;
            d.box_nb[t] = p.box_nb[t];
        }
    }
    d.stats = h->d_stats; d.log = h->d_log;
    d.op_out_scalar = h->d_scalar;
    d.op_repeat = 1;
    d.replicate_stale_sr = h->opt.replicate_stale_sr; d.max_outer = h->opt.max_outer; d.max_inner = h->opt.max_inner;
    return XM_OK;
}

static int ensure_io(xm_handle* h, int r) {
    const int rcap = (h->world > 1) ? std::max(r, h->comm_maxr) : r;      // see carve(): never re-allocate inside a collective
    const size_t bR = (size_t)h->n3 * rcap * sizeof(double), bS = (size_t)h->N * sizeof(double);
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

// Kernel instantiations live in xm_inst.cu, compiled 18 times (groups of padded ranks x Q.Y path x one-GPU / communicator
// build) so the build parallelises.
#define XM_DECL(g, p, m) cudaError_t xm_launch_group##g##_path##p##_mg##m(int kind, int RP, const xm_handle* h, const Dev& d, int opcode, size_t dyn, cudaStream_t st);
#define XM_DECL3(g, m) XM_DECL(g, 0, m) XM_DECL(g, 1, m) XM_DECL(g, 2, m)
XM_DECL3(0, 0) XM_DECL3(1, 0) XM_DECL3(2, 0) XM_DECL3(0, 1) XM_DECL3(1, 1) XM_DECL3(2, 1)
#undef XM_DECL3
#undef XM_DECL

static cudaError_t launch_any(int kind, const xm_handle* h, const Dev& d, int opcode, const Plan& p, cudaStream_t st) {
    typedef cudaError_t (*Fn)(int, int, const xm_handle*, const Dev&, int, size_t, cudaStream_t);
    static const Fn table[2][3][3] = {   // [communicator?][rank group: RP 3,4,5 / 6,8,10 / 12,16,20][path: TMA ring / dense direct / block-CSR]
        {{xm_launch_group0_path0_mg0, xm_launch_group0_path1_mg0, xm_launch_group0_path2_mg0},
         {xm_launch_group1_path0_mg0, xm_launch_group1_path1_mg0, xm_launch_group1_path2_mg0},
         {xm_launch_group2_path0_mg0, xm_launch_group2_path1_mg0, xm_launch_group2_path2_mg0}},
        {{xm_launch_group0_path0_mg1, xm_launch_group0_path1_mg1, xm_launch_group0_path2_mg1},
         {xm_launch_group1_path0_mg1, xm_launch_group1_path1_mg1, xm_launch_group1_path2_mg1},
         {xm_launch_group2_path0_mg1, xm_launch_group2_path1_mg1, xm_launch_group2_path2_mg1}}};
    const int g = (p.RP <= 5) ? 0 : (p.RP <= 10) ? 1 : 2;
    const int path = d.use_tma ? 0 : (d.Q ? 1 : 2);
    return table[d.world > 1 ? 1 : 0][g][path](kind, p.RP, h, d, opcode, p.dyn_smem, st);
}
static cudaError_t launch_solve(const xm_handle* h, const Dev& d, const Plan& p, cudaStream_t st) { return launch_any(0, h, d, 0, p, st); }
static cudaError_t launch_ops(const xm_handle* h, const Dev& d, int opcode, const Plan& p, cudaStream_t st) { return launch_any(1, h, d, opcode, p, st); }

static int prepare(xm_handle* h, int r, Plan* plan) {
    if (!h) return XM_EINVAL;
    if (r < 3 || r > XM_MAX_RANK) { h->err = "rank out of range [3,20]"; return XM_EINVAL; }
    if (h->N <= 0 || (!h->is_bsr && !h->Qp)) { h->err = "no Q set"; return XM_EINVAL; }
    if (h->world > 1) {
        if (!h->comm_connected) { h->err = "communicator not connected (xm_comm_connect)"; return XM_EINVAL; }
        if (h->comm_broken) { h->err = "communicator desynchronised by an earlier abort (xm_comm_reset on every rank)"; return XM_ESYNC; }
        if (h->N != h->comm_N || r > h->comm_maxr) { h->err = "problem does not fit the communicator (camera count / max rank)"; return XM_EINVAL; }
    }
    XM_CUDA(h, cudaSetDevice(h->device));
    *plan = make_plan(h, r);
    int rc = carve(h, r, *plan);
    if (rc) return rc;
    rc = ensure_io(h, r);
    if (rc) { h->err = "io staging alloc failed"; return rc; }
    if (h->world == 1) {      // with a communicator the counters are monotone across launches: a peer may already be arriving
        XM_CUDA(h, cudaMemsetAsync(h->d_bar, 0, 256, h->stream));
        XM_CUDA(h, cudaMemsetAsync(h->d_abort, 0, 256, h->stream));
    }
    return XM_OK;
}

static int check_abort(xm_handle* h) {
    int ab = 0;
    const int* flag = h->world > 1 ? (const int*)(h->arena + h->off_abort) : h->d_abort;
    XM_CUDA(h, cudaStreamSynchronize(h->stream));
    XM_CUDA(h, cudaMemcpyAsync(&ab, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    XM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (ab) { h->err = "device grid barrier timed out"; if (h->world > 1) h->comm_broken = true; return XM_ESYNC; }
    return XM_OK;
}

// Where a launch leaves its results on THIS device.  Single GPU: straight into the caller's / staging buffers.  With a
// communicator every CTA stores its cameras into every rank's arena copy; the caller's buffer is filled from the local copy.
struct OutBind { double* R; double* s; };
static OutBind bind_out(xm_handle* h, Dev& d, double* R_dev, double* s_dev) {
    if (h->world == 1) { d.outR_peer[0] = R_dev; d.outS_peer[0] = s_dev; return {R_dev, s_dev}; }
    for (int w = 0; w < h->world; ++w) {
        d.outR_peer[w] = (double*)(h->peer_arena[w] + h->off_outR); d.outS_peer[w] = (double*)(h->peer_arena[w] + h->off_outS);
    }
    return {(double*)(h->arena + h->off_outR), (double*)(h->arena + h->off_outS)};
}

// ------------------------------------------------------------------------------------------------ Q.Y
static int qy_common(xm_handle* h, int r, double alpha, const double* X, double* out, bool dev_ptrs) {
    XmRange nvtx_range("xm_qy");
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!X || !out) return XM_EINVAL;
    Dev d = h->dev;
    const cudaMemcpyKind kin = dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (d.x_cam_major) {      // block-CSR: camera-major operand
        const double* src = X;
        if (!dev_ptrs) {
            XM_CUDA(h, cudaMemcpyAsync(h->io_P, X, (size_t)d.n3 * r * sizeof(double), kin, h->stream));
            src = h->io_P;
        }
        const long long tot = (long long)d.n3 * r;
        xm_operand_cam_major_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(src, d.n3, r, d.Xt);
        XM_CUDA(h, cudaGetLastError());
        h->launches++;
    } else {
        // the wire layout (3N x r column-major) IS the operand layout up to the padded leading dimension
        XM_CUDA(h, cudaMemcpy2DAsync(d.Xt, (size_t)d.ldq * sizeof(double), X, (size_t)d.n3 * sizeof(double),
                                     (size_t)d.n3 * sizeof(double), r, kin, h->stream));
    }
    d.qy_alpha = alpha;
    const OutBind ob = bind_out(h, d, dev_ptrs ? out : h->io_Rout, h->io_sout);
    XM_CUDA(h, launch_ops(h, d, 0, p, h->stream));
    h->launches++;
    if (!dev_ptrs) XM_CUDA(h, cudaStreamSynchronize(h->stream));      // see tr_common: never block inside a pageable copy behind a kernel
    if (!dev_ptrs || ob.R != out)
        XM_CUDA(h, cudaMemcpyAsync(out, ob.R, (size_t)d.n3 * r * sizeof(double), dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    if (!dev_ptrs || h->world > 1) return check_abort(h);     // synchronises
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
    if (!avg_ms || iters == 0) return XM_EINVAL;
    Dev d = h->dev;
    d.qy_alpha = 1.0;
    (void)bind_out(h, d, h->io_Rout, h->io_sout);
    cudaEvent_t e0, e1;
    XM_CUDA(h, cudaEventCreate(&e0)); XM_CUDA(h, cudaEventCreate(&e1));
    // iters > 0: `iters` products inside ONE launch (steady state, ring prefetch across products, like the solver);
    d.op_repeat = iters;        // negative: |iters| products with a grid barrier after each (the solver's lock-step)
    if (iters < 0) iters = -iters;
    XM_CUDA(h, launch_ops(h, d, 0, p, h->stream));            // warm-up launch
    XM_CUDA(h, cudaEventRecord(e0, h->stream));
    XM_CUDA(h, launch_ops(h, d, 0, p, h->stream));
    XM_CUDA(h, cudaEventRecord(e1, h->stream));
    XM_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    XM_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->launches += 2;
    *avg_ms = (double)ms / iters;
    return check_abort(h);
}

// measurement hook: average device time (us) of one grid barrier of the persistent kernel (|iters| barriers, one launch)
extern "C" int xm_bench_barrier(xm_handle* h, int r, int iters, double* avg_us) {
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!avg_us || iters == 0) return XM_EINVAL;
    Dev d = h->dev;
    d.op_repeat = iters;          // negative: |iters| barriers followed by |iters| operand exchanges (push, unpack, operand_sync)
    if (iters < 0) iters = -iters;
    cudaEvent_t e0, e1;
    XM_CUDA(h, cudaEventCreate(&e0)); XM_CUDA(h, cudaEventCreate(&e1));
    (void)bind_out(h, d, h->io_Rout, h->io_sout);
    XM_CUDA(h, launch_ops(h, d, 5, p, h->stream));
    if (h->world == 1) XM_CUDA(h, cudaMemsetAsync(h->d_bar, 0, 256, h->stream));
    XM_CUDA(h, cudaEventRecord(e0, h->stream));
    XM_CUDA(h, launch_ops(h, d, 5, p, h->stream));
    XM_CUDA(h, cudaEventRecord(e1, h->stream));
    XM_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    XM_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->launches += 2;
    *avg_us = (double)ms * 1e3 / iters;
    return check_abort(h);
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
    XmRange nvtx_range("xm_trust_region");
    Plan p;
    int rc = prepare(h, r, &p);
    if (rc) return rc;
    if (!R0 || !s0 || !gradtol_inout || !R_out || !s_out || (ls_step != 0.0 && !v)) { h->err = "null argument"; return XM_EINVAL; }
    Dev d = h->dev;
    const size_t bR = (size_t)d.n3 * r * sizeof(double), bS = (size_t)d.N * sizeof(double);
    OutBind ob;
    if (dev_ptrs) {
        d.R0 = R0; d.s0 = s0; d.vdir = v;
        ob = bind_out(h, d, R_out, s_out);
    } else {
        XM_CUDA(h, cudaMemcpyAsync(h->io_R0, R0, bR, cudaMemcpyHostToDevice, h->stream));
        XM_CUDA(h, cudaMemcpyAsync(h->io_s0, s0, bS, cudaMemcpyHostToDevice, h->stream));
        if (ls_step != 0.0) XM_CUDA(h, cudaMemcpyAsync(h->io_v, v, 3 * bS, cudaMemcpyHostToDevice, h->stream));
        d.R0 = h->io_R0; d.s0 = h->io_s0; d.vdir = h->io_v;
        ob = bind_out(h, d, h->io_Rout, h->io_sout);
    }
    d.lam = lam; d.gradtol = *gradtol_inout; d.ls_step = ls_step; d.max_time = max_time;
    cudaEvent_t e0, e1;
    XM_CUDA(h, cudaEventCreate(&e0)); XM_CUDA(h, cudaEventCreate(&e1));
    XM_CUDA(h, cudaEventRecord(e0, h->stream));
    XM_CUDA(h, launch_solve(h, d, p, h->stream));
    XM_CUDA(h, cudaEventRecord(e1, h->stream));
    h->launches++;
    // Host (pageable) destinations: wait for the kernel with a plain stream synchronisation FIRST.  A pageable device-to-host copy
    // blocks inside the driver until the kernel ahead of it has finished; when two members of a communicator share one device
    // (one context: the loop-back team of the single-GPU test suite) that blocked call keeps the peer's thread from launching its
    // own kernel — the two persistent kernels would wait for each other until the watchdog fires.
    if (!dev_ptrs) XM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!dev_ptrs || ob.R != R_out) {
        const cudaMemcpyKind kout = dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        XM_CUDA(h, cudaMemcpyAsync(R_out, ob.R, bR, kout, h->stream));
        XM_CUDA(h, cudaMemcpyAsync(s_out, ob.s, bS, kout, h->stream));
    }
    XM_CUDA(h, cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, h->stream));
    XM_CUDA(h, cudaMemcpyAsync(h->h_log, h->d_log, sizeof(LogRec) * kLogCap, cudaMemcpyDeviceToHost, h->stream));
    rc = check_abort(h);      // synchronises the stream
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc) return rc;
    const DevStats& S = *h->h_stats;
    if (S.aborted) { h->err = "solver kernel aborted"; if (h->world > 1) h->comm_broken = true; return XM_ESYNC; }
    *gradtol_inout = S.gradtol_out;
    if (primal_out) *primal_out = S.primal;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->exit_code = S.exit_code; stats->outer_iters = S.outer_iters; stats->tcg_iters = S.tcg_iters;
        stats->qy_products = S.qy_products; stats->n_log = S.n_log; stats->primal = S.primal; stats->gradnorm = S.gradnorm;
        stats->solve_ms = ms; stats->qy_ms = S.qy_ns * 1e-6; stats->sync_ms = S.sync_ns * 1e-6;
        stats->grid_ctas = p.G; stats->threads_per_cta = p.NT; stats->ksplit = p.KS; stats->launches = 1;
        for (int q = 0; q < 4; ++q) stats->phase_ms[q] = S.dbg[q] * 1e-6;
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
    XmRange nvtx_range("xm_op");
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
    const OutBind ob = bind_out(h, d, h->io_Rout, h->io_sout);
    XM_CUDA(h, launch_ops(h, d, opcode, p, h->stream));
    h->launches++;
    XM_CUDA(h, cudaStreamSynchronize(h->stream));                      // see tr_common
    if (outR) XM_CUDA(h, cudaMemcpyAsync(outR, ob.R, bR, cudaMemcpyDeviceToHost, h->stream));
    if (outS) XM_CUDA(h, cudaMemcpyAsync(outS, ob.s, bS, cudaMemcpyDeviceToHost, h->stream));
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

// 3x3 diagonal blocks of the operator (N x 9 doubles, row-major per block) into a device buffer; collective on a communicator
extern "C" int xm_op_diag_blocks_dev(xm_handle* h, double* out9N_dev) {
    Plan p;
    int rc = prepare(h, 3, &p);
    if (rc) return rc;
    if (!out9N_dev) return XM_EINVAL;
    Dev d = h->dev;
    const OutBind ob = bind_out(h, d, out9N_dev, h->io_sout);        // 9 N doubles fit the 3N x r (r >= 3) result copy of a communicator
    XM_CUDA(h, launch_ops(h, d, 6, p, h->stream));
    h->launches++;
    if (ob.R != out9N_dev)
        XM_CUDA(h, cudaMemcpyAsync(out9N_dev, ob.R, (size_t)h->N * 9 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return check_abort(h);
}
extern "C" int xm_op_diag_blocks(xm_handle* h, double* out9N) {
    if (!h || !out9N || h->N <= 0) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    double* tmp = nullptr;
    XM_CUDA(h, cudaMalloc(&tmp, (size_t)h->N * 9 * sizeof(double)));
    int rc = xm_op_diag_blocks_dev(h, tmp);
    if (rc == XM_OK && cudaMemcpy(out9N, tmp, (size_t)h->N * 9 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); rc = XM_ECUDA; }
    cudaFree(tmp);
    return rc;
}

// debug: the 8 raw counters DevStats.dbg left on the device by the last launch (ns; meaning depends on the kernel/opcode)
extern "C" int xm_debug_counters(xm_handle* h, unsigned long long* out8) {
    if (!h || !out8) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    XM_CUDA(h, cudaMemcpy(out8, (const char*)h->d_stats + offsetof(DevStats, dbg), 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return XM_OK;
}

// debug: (tag, ns) pairs recorded by the last profiled solve (opt.profile = 1); out must hold 256 entries
extern "C" int xm_debug_trace(xm_handle* h, unsigned long long* out) {
    if (!h || !out) return XM_EINVAL;
    memcpy(out, h->h_stats->trace, sizeof(h->h_stats->trace));
    return XM_OK;
}

extern "C" int xm_escape_scale(int n, double* v, const double* s) {
    if (!v || !s || n < 0) return XM_EINVAL;
    for (int i = 0; i < n; ++i) { v[3 * i] /= s[i]; v[3 * i + 1] /= s[i]; v[3 * i + 2] /= s[i]; }   // XM_main.cu:8-16
    return XM_OK;
}
