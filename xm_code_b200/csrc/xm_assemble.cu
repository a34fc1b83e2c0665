// xm_assemble.cu — Q assembly on the device behind xm_create_matrix (replaces utils/creatematrix.py:52-341).
//
// Observations (camera i, landmark k, weight w, camera-frame point p~) of  sum w || s_i R_i p~_ik + t_i - p_k ||^2 ; eliminating
// translations and landmarks with t_1 = 0 gives (same algebra as the reference, different route — SURVEY.md §3.4):
//     Q = Q1 - Vl Dl^-1 Vl^T - Bb Scb^-1 Bb^T ,   Q1 = blkdiag_i(sum_k w p~ p~^T),
//     Sc = diag(dc) - W Dl^-1 W^T  (reduced camera Laplacian; Scb: camera 1's row/column removed),
//     B  = Vc + Vl Dl^-1 W^T       (3N x N; Bb: camera 1's column removed).
// The reference forms the normal equations densely in M (V3_bar_F.toarray(), an (N+M) x 3N dense block solve on the host).  Here the
// landmark block (diagonal) is eliminated first by ONE kernel over the co-observation pairs of every landmark (a warp per
// landmark, FP64 atomics into Q / Sc / B), leaving an (N-1) x (N-1) Cholesky (cuSOLVER), one triangular solve and one SYRK (cuBLAS —
// plain library GEMM-class calls): BAL-Final-sized Q (13 682 cameras, 13.5 GB) assembles in seconds.  The result is written
// straight into the handle's padded operator buffer (Q is symmetric: row-major == column-major), so a solve can follow without
// any host round trip.  Abar (the map back to translations / landmarks, creatematrix.py:308-311) is optional: it is dense in
// M x 3N and only makes sense for small problems.
#include "xm_host.h"
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <vector>
#include <string>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <chrono>

int xm_internal_libs(xm_handle* h, void** cublas_out, void** cusolver_out);    // xm_capi.cu

namespace {

struct Obs { const int* cam; const int* lm; const double* w; const double* pt; };      // device arrays, n_obs entries (pt: n_obs x 3)

// per landmark: dl[k] = sum of its weights (serial over its few observations: fixed order)
__global__ void landmark_degree_kernel(Obs o, const int* __restrict__ lm_ptr, const int* __restrict__ lm_obs, int M, double* __restrict__ dl) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    double acc = 0;
    for (int e = lm_ptr[k]; e < lm_ptr[k + 1]; ++e) acc += o.w[lm_obs[e]];
    dl[k] = acc;
}
// per camera: dc = sum w ; Q[3i+a, 3i+b] = sum w p~ p~^T (Q1) ; B[3i+a, i] = sum w p~ (Vc) ; Sc[i, i] = dc
__global__ void camera_terms_kernel(Obs o, const int* __restrict__ cam_ptr, const int* __restrict__ cam_obs, int N, int n3, int ldq,
                                    double* __restrict__ Q, double* __restrict__ B, double* __restrict__ Sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double dc = 0, v[3] = {0, 0, 0}, q[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int e = cam_ptr[i]; e < cam_ptr[i + 1]; ++e) {
        const int ob = cam_obs[e];
        const double w = o.w[ob];
        const double p[3] = {o.pt[(size_t)ob * 3], o.pt[(size_t)ob * 3 + 1], o.pt[(size_t)ob * 3 + 2]};
        dc += w;
        for (int a = 0; a < 3; ++a) { v[a] += w * p[a]; for (int b = 0; b < 3; ++b) q[a][b] += w * p[a] * p[b]; }
    }
    for (int a = 0; a < 3; ++a) {
        atomicAdd(&B[(size_t)i * n3 + 3 * i + a], v[a]);
        for (int b = 0; b < 3; ++b) atomicAdd(&Q[(size_t)(3 * i + b) * ldq + 3 * i + a], q[a][b]);
    }
    atomicAdd(&Sc[(size_t)i * N + i], dc);
}
// one warp per landmark: all ordered pairs (x, y) of its observations
//   Q[3 cx + a, 3 cy + b] -= (w p~)_x[a] (w p~)_y[b] / dl ;  Sc[cx, cy] -= w_x w_y / dl ;  B[3 cx + a, cy] -= (w p~)_x[a] w_y / dl
__global__ void landmark_pairs_kernel(Obs o, const int* __restrict__ lm_ptr, const int* __restrict__ lm_obs, const double* __restrict__ dl,
                                      int M, int N, int n3, int ldq, double* __restrict__ Q, double* __restrict__ B, double* __restrict__ Sc) {
    const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= M) return;
    const int e0 = lm_ptr[warp], d = lm_ptr[warp + 1] - e0;
    const double inv = 1.0 / dl[warp];
    for (long long p = lane; p < (long long)d * d; p += 32) {
        const int x = lm_obs[e0 + (int)(p / d)], y = lm_obs[e0 + (int)(p % d)];
        const int cx = o.cam[x], cy = o.cam[y];
        const double wx = o.w[x], wy = o.w[y];
        const double px[3] = {wx * o.pt[(size_t)x * 3], wx * o.pt[(size_t)x * 3 + 1], wx * o.pt[(size_t)x * 3 + 2]};
        const double py[3] = {wy * o.pt[(size_t)y * 3], wy * o.pt[(size_t)y * 3 + 1], wy * o.pt[(size_t)y * 3 + 2]};
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) atomicAdd(&Q[(size_t)(3 * cy + b) * ldq + 3 * cx + a], -px[a] * py[b] * inv);
            atomicAdd(&B[(size_t)cy * n3 + 3 * cx + a], -px[a] * wy * inv);
        }
        atomicAdd(&Sc[(size_t)cy * N + cx], -wx * wy * inv);
    }
}
// upper triangle <- lower triangle (column-major view with leading dimension ld; SYRK only updated the lower one)
__global__ void mirror_lower_kernel(double* __restrict__ Q, int n3, int ld) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    if (bj > bi) return;                                          // tiles on or below the diagonal are the sources
    const int i = bi + threadIdx.x;                               // row (fast index of the column-major view)
    for (int t = threadIdx.y; t < 32; t += blockDim.y) {
        const int j = bj + t;
        tile[t][threadIdx.x] = (i < n3 && j < n3) ? Q[(size_t)j * ld + i] : 0.0;     // tile[j - bj][i - bi] = Q[i, j]
    }
    __syncthreads();
    for (int t = threadIdx.y; t < 32; t += blockDim.y) {
        const int ii = bi + t, jj = bj + threadIdx.x;             // write Q[jj, ii] = Q[ii, jj] for ii > jj
        if (ii < n3 && jj < n3 && ii > jj) Q[(size_t)ii * ld + jj] = tile[threadIdx.x][t];
    }
}
// Abar rows 0 .. N-2:  a_t[c, j] = -Zt[j, c]   (Zt = Bb Scb^-1, n3 x (N-1) column-major)
__global__ void abar_translations_kernel(const double* __restrict__ Zt, int N, int n3, long long rows, double* __restrict__ Abar) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n3 * (N - 1)) return;
    const int j = (int)(t % n3), c = (int)(t / n3);
    Abar[(size_t)j * rows + c] = -Zt[(size_t)c * n3 + j];
}
// Abar rows N-1 ..:  b_p[k, j] = ( sum_{obs of k} [ j in the obs' camera rows ] w p~  +  sum_{obs of k, cam c >= 1} w a_t[c-1, j] ) / dl_k
__global__ void abar_landmarks_kernel(Obs o, const int* __restrict__ lm_ptr, const int* __restrict__ lm_obs, const double* __restrict__ dl,
                                      const double* __restrict__ Zt, int N, int n3, long long rows, double* __restrict__ Abar) {
    const int k = blockIdx.x;
    const int e0 = lm_ptr[k], e1 = lm_ptr[k + 1];
    const double inv = 1.0 / dl[k];
    for (int j = threadIdx.x; j < n3; j += blockDim.x) {
        double acc = 0;
        for (int e = e0; e < e1; ++e) {
            const int ob = lm_obs[e], c = o.cam[ob];
            const double w = o.w[ob];
            if (j / 3 == c) acc += w * o.pt[(size_t)ob * 3 + (j - 3 * c)];
            if (c >= 1) acc -= w * Zt[(size_t)(c - 1) * n3 + j];
        }
        Abar[(size_t)j * rows + (N - 1) + k] = acc * inv;
    }
}

struct Bufs {
    std::vector<void*> p;
    ~Bufs() { for (void* q : p) cudaFree(q); }
    template <class T> T* get(size_t n) { void* q = nullptr; if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return nullptr; } p.push_back(q); return (T*)q; }
};
struct Libs { cublasHandle_t cb = nullptr; cusolverDnHandle_t cs = nullptr; };      // views of the handle's cached library handles

// group the observation indices by key (counting sort: stable, original order inside a group)
void group_by(const int* key, int64_t n, int nkeys, std::vector<int>& ptr, std::vector<int>& idx) {
    ptr.assign((size_t)nkeys + 1, 0);
    for (int64_t e = 0; e < n; ++e) ptr[(size_t)key[e] + 1]++;
    for (int k = 0; k < nkeys; ++k) ptr[k + 1] += ptr[k];
    idx.resize((size_t)n);
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < n; ++e) idx[(size_t)fill[key[e]]++] = (int)e;
}

}  // namespace

#define ASM_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); cudaGetLastError(); return XM_ECUDA; } } while (0)

extern "C" int xm_create_matrix(xm_handle* h, int n_cameras, int n_landmarks, int64_t n_obs, const int* cam, const int* lm,
                                const double* w, const double* pts, double* Q_out, double* Abar_out, double* assemble_ms_out) {
    XmRange nvtx_range("xm_create_matrix");
    if (!h || !cam || !lm || !w || !pts || n_cameras < 2 || n_landmarks < 1 || n_obs < 1 || n_obs > 2000000000LL) return XM_EINVAL;
    if (h->world > 1) { h->err = "xm_create_matrix assembles on one GPU (attach the communicator afterwards and upload the slabs)"; return XM_EUNSUPPORTED; }
    const int N = n_cameras, M = n_landmarks, n3 = 3 * N, ldq = (n3 + 63) / 64 * 64;
    for (int64_t e = 0; e < n_obs; ++e)
        if (cam[e] < 0 || cam[e] >= N || lm[e] < 0 || lm[e] >= M) { h->err = "observation with a camera / landmark index out of range"; return XM_EINVAL; }
    std::vector<int> lm_ptr, lm_obs, cam_ptr, cam_obs;
    group_by(lm, n_obs, M, lm_ptr, lm_obs);
    group_by(cam, n_obs, N, cam_ptr, cam_obs);
    for (int k = 0; k < M; ++k) if (lm_ptr[k + 1] == lm_ptr[k]) { h->err = "every landmark needs at least one observation"; return XM_EINVAL; }
    const auto t_wall0 = std::chrono::steady_clock::now();
    ASM_CUDA(cudaSetDevice(h->device));
    // the operator buffer of the handle receives the result
    if (h->Qp_cap < (size_t)n3 * ldq * sizeof(double) || !h->Qp) {
        if (h->Qp) cudaFree(h->Qp);
        h->Qp = nullptr; h->Qp_cap = 0;
        if (cudaMalloc(&h->Qp, (size_t)n3 * ldq * sizeof(double)) != cudaSuccess) { cudaGetLastError(); h->err = "cudaMalloc of Q failed"; return XM_ENOMEM; }
        h->Qp_cap = (size_t)n3 * ldq * sizeof(double);
    }
    h->N = 0; h->n3 = 0;                                           // not a valid operator until the assembly has succeeded
    Bufs mem;
    Libs lib;
    int* d_cam = mem.get<int>(n_obs); int* d_lm = mem.get<int>(n_obs); double* d_w = mem.get<double>(n_obs); double* d_pt = mem.get<double>(3 * (size_t)n_obs);
    int* d_lm_ptr = mem.get<int>((size_t)M + 1); int* d_lm_obs = mem.get<int>(n_obs); int* d_cam_ptr = mem.get<int>((size_t)N + 1); int* d_cam_obs = mem.get<int>(n_obs);
    double* dl = mem.get<double>(M); double* Sc = mem.get<double>((size_t)N * N); double* B = mem.get<double>((size_t)n3 * N); int* info = mem.get<int>(1);
    if (!d_cam || !d_lm || !d_w || !d_pt || !d_lm_ptr || !d_lm_obs || !d_cam_ptr || !d_cam_obs || !dl || !Sc || !B || !info) { h->err = "assembly workspace cudaMalloc failed"; return XM_ENOMEM; }
    cudaStream_t st = h->stream;
    struct Events { cudaEvent_t a = nullptr, b = nullptr; ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evs;    // freed on every exit path
    ASM_CUDA(cudaEventCreate(&evs.a)); ASM_CUDA(cudaEventCreate(&evs.b));
    cudaEvent_t ev0 = evs.a, ev1 = evs.b;
    ASM_CUDA(cudaEventRecord(ev0, st));
    ASM_CUDA(cudaMemcpyAsync(d_cam, cam, n_obs * sizeof(int), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_lm, lm, n_obs * sizeof(int), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_w, w, n_obs * sizeof(double), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_pt, pts, 3 * n_obs * sizeof(double), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_lm_ptr, lm_ptr.data(), ((size_t)M + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_lm_obs, lm_obs.data(), n_obs * sizeof(int), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_cam_ptr, cam_ptr.data(), ((size_t)N + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemcpyAsync(d_cam_obs, cam_obs.data(), n_obs * sizeof(int), cudaMemcpyHostToDevice, st));
    ASM_CUDA(cudaMemsetAsync(h->Qp, 0, (size_t)n3 * ldq * sizeof(double), st));
    ASM_CUDA(cudaMemsetAsync(Sc, 0, (size_t)N * N * sizeof(double), st));
    ASM_CUDA(cudaMemsetAsync(B, 0, (size_t)n3 * N * sizeof(double), st));
    const bool trace = getenv("XM_ASM_TRACE") != nullptr;
    auto mark = [&](const char* what) { if (trace) { cudaStreamSynchronize(st); fprintf(stderr, "[xm_create_matrix] %-28s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wall0).count()); } };
    mark("uploads + memsets");
    const Obs o{d_cam, d_lm, d_w, d_pt};
    landmark_degree_kernel<<<(M + 255) / 256, 256, 0, st>>>(o, d_lm_ptr, d_lm_obs, M, dl);
    camera_terms_kernel<<<(N + 127) / 128, 128, 0, st>>>(o, d_cam_ptr, d_cam_obs, N, n3, ldq, h->Qp, B, Sc);
    landmark_pairs_kernel<<<(unsigned)(((long long)M * 32 + 255) / 256), 256, 0, st>>>(o, d_lm_ptr, d_lm_obs, dl, M, N, n3, ldq, h->Qp, B, Sc);
    ASM_CUDA(cudaGetLastError());
    h->launches += 3;
    mark("elimination kernels");
    // (N-1) x (N-1) Cholesky of the reduced camera Laplacian, Y = Bb L^-T, Q -= Y Y^T (lower triangle), mirror
    { const int lrc = xm_internal_libs(h, (void**)&lib.cb, (void**)&lib.cs); if (lrc) return lrc; }
    double* Scb = Sc + (size_t)N + 1;                              // Sc[1:, 1:], leading dimension N
    double* Bb = B + (size_t)n3;                                   // B[:, 1:]
    int lwork = 0;
    if (cusolverDnDpotrf_bufferSize(lib.cs, CUBLAS_FILL_MODE_LOWER, N - 1, Scb, N, &lwork) != CUSOLVER_STATUS_SUCCESS) { h->err = "potrf_bufferSize"; return XM_ECUDA; }
    double* work = mem.get<double>((size_t)std::max(lwork, 1));
    if (!work) { h->err = "potrf workspace cudaMalloc failed"; return XM_ENOMEM; }
    if (cusolverDnDpotrf(lib.cs, CUBLAS_FILL_MODE_LOWER, N - 1, Scb, N, work, lwork, info) != CUSOLVER_STATUS_SUCCESS) { h->err = "cusolverDnDpotrf"; return XM_ECUDA; }
    int hinfo = 0;
    ASM_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
    ASM_CUDA(cudaStreamSynchronize(st));
    mark("potrf");
    if (hinfo != 0) { h->err = "the reduced camera Laplacian is not positive definite (the camera-landmark graph is disconnected?)"; return XM_EINVAL; }
    const double one = 1.0, mone = -1.0;
    if (cublasDtrsm(lib.cb, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n3, N - 1, &one, Scb, N, Bb, n3) != CUBLAS_STATUS_SUCCESS) { h->err = "cublasDtrsm"; return XM_ECUDA; }
    if (cublasDsyrk(lib.cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n3, N - 1, &mone, Bb, n3, &one, h->Qp, ldq) != CUBLAS_STATUS_SUCCESS) { h->err = "cublasDsyrk"; return XM_ECUDA; }
    mirror_lower_kernel<<<dim3((n3 + 31) / 32, (n3 + 31) / 32), dim3(32, 8), 0, st>>>(h->Qp, n3, ldq);
    ASM_CUDA(cudaGetLastError());
    h->launches += 1;
    ASM_CUDA(cudaEventRecord(ev1, st));
    mark("trsm + syrk + mirror");
    if (Abar_out) {
        // Zt = Y L^-1 = Bb Scb^-1 ; a_t = -Zt^T ; b_p from the observations and a_t
        const long long rows = (long long)N + M - 1;
        double* Abar = mem.get<double>((size_t)rows * n3);
        if (!Abar) { h->err = "Abar is dense in (N + M - 1) x 3N: cudaMalloc failed (pass NULL for large problems)"; return XM_ENOMEM; }
        if (cublasDtrsm(lib.cb, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n3, N - 1, &one, Scb, N, Bb, n3) != CUBLAS_STATUS_SUCCESS) { h->err = "cublasDtrsm"; return XM_ECUDA; }
        abar_translations_kernel<<<(unsigned)(((long long)n3 * (N - 1) + 255) / 256), 256, 0, st>>>(Bb, N, n3, rows, Abar);
        abar_landmarks_kernel<<<M, 256, 0, st>>>(o, d_lm_ptr, d_lm_obs, dl, Bb, N, n3, rows, Abar);
        ASM_CUDA(cudaGetLastError());
        h->launches += 2;
        ASM_CUDA(cudaMemcpyAsync(Abar_out, Abar, (size_t)rows * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (Q_out)
        ASM_CUDA(cudaMemcpy2DAsync(Q_out, (size_t)n3 * sizeof(double), h->Qp, (size_t)ldq * sizeof(double), (size_t)n3 * sizeof(double), n3, cudaMemcpyDeviceToHost, st));
    ASM_CUDA(cudaStreamSynchronize(st));
    mark("Abar + copies out");
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    if (assemble_ms_out) *assemble_ms_out = ms;
    h->n3 = n3; h->N = N; h->ldq = ldq; h->is_bsr = false; h->cam0 = 0; h->cam1 = N;
    return XM_OK;
}
