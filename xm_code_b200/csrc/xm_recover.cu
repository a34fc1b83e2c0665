// xm_recover.cu — solution recovery behind xm_recover (replaces recover_XM, utils/recoversolution.py:4-86).
//
// Same result as the reference, different mechanics:
//   * rank r > 3 -> 3: the reference eigendecomposes the 3N x 3N matrix (sR)(sR)^T on the host (:11-23).  Its top-3
//     factor V_3 sqrt(L_3) equals sR W_3 with W_3 the top-3 eigenvectors of the r x r Gram matrix (sR)^T (sR) — the same
//     matrix up to a sign per column, which the anchoring step removes.  Here: a deterministic two-stage grid reduction of
//     the Gram matrix, then a cyclic Jacobi eigensolve of the r x r matrix (r <= 20) by one thread.
//   * per camera: 3 x 3 block, scale = ||block||_F / sqrt(3) (:42-44), anchoring by camera 0 (:47-48), projection to O(3)
//     by the polar factor U V^T (:50-73) from a one-sided Jacobi SVD held in registers — one thread per camera (a 3 x 3
//     problem does not fill a warp; the work is ~300 flops per camera).
//   * the reference counts cameras with det(U V^T) < 0 and negates everything if they are the majority (:62-63) before
//     projecting; polar(-M) = -polar(M), so the projection runs once and the sign is applied afterwards.
//   * [t p] = Abar (s R)^T (:76-86): Abar is (N + M - 1) x 3N column-major — one thread per row streams it fully coalesced
//     (HBM-bound: 8 B per element, 6 flops), columns split over blockIdx.y with a fixed-order second pass.
#include "xm_host.h"
#include <vector>
#include <cmath>
#include <algorithm>

namespace {

constexpr int kMaxR = XM_MAX_RANK;

// ---- stage 1: per-CTA partial Gram matrices  G[j][k] = sum_rows sR[row][j] sR[row][k]   (R: 3N x r column-major)
__global__ void gram_partial_kernel(const double* __restrict__ R, const double* __restrict__ s, int n3, int r, double* __restrict__ part) {
    extern __shared__ double tile[];                       // [64][r]
    const int npair = r * r;
    const int rows_per_cta = (n3 + gridDim.x - 1) / gridDim.x;
    const int row_lo = blockIdx.x * rows_per_cta, row_hi = min(n3, row_lo + rows_per_cta);
    const int pj = threadIdx.x / r, pk = threadIdx.x % r;
    double acc = 0.0;
    for (int base = row_lo; base < row_hi; base += 64) {
        const int nr = min(64, row_hi - base);
        for (int t = threadIdx.x; t < nr * r; t += blockDim.x) {
            const int j = t / nr, rr = t % nr;             // consecutive threads along rows: coalesced column-major reads
            tile[rr * r + j] = R[(size_t)j * n3 + base + rr] * s[(base + rr) / 3];
        }
        __syncthreads();
        if (threadIdx.x < npair)
            for (int rr = 0; rr < nr; ++rr) acc = fma(tile[rr * r + pj], tile[rr * r + pk], acc);
        __syncthreads();
    }
    if (threadIdx.x < npair) part[(size_t)blockIdx.x * npair + threadIdx.x] = acc;
}

// ---- stage 2: fixed-order sum of the partials, cyclic Jacobi, top-3 eigenvectors (descending eigenvalues)
__global__ void gram_eig_kernel(const double* __restrict__ part, int nparts, int r, double* __restrict__ W3, double* __restrict__ eig) {
    __shared__ double A[kMaxR * kMaxR], V[kMaxR * kMaxR];
    for (int t = threadIdx.x; t < r * r; t += blockDim.x) {
        double a = 0.0;
        for (int p = 0; p < nparts; ++p) a += part[(size_t)p * r * r + t];
        A[t] = a; V[t] = (t / r == t % r) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int j = 0; j < r; ++j) for (int k = j + 1; k < r; ++k) { const double m = 0.5 * (A[j * r + k] + A[k * r + j]); A[j * r + k] = A[k * r + j] = m; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < r; ++j) { diag += A[j * r + j] * A[j * r + j]; for (int k = j + 1; k < r; ++k) off += A[j * r + k] * A[j * r + k]; }
        if (off <= 1e-32 * diag) break;
        for (int p = 0; p < r; ++p)
            for (int q = p + 1; q < r; ++q) {
                const double apq = A[p * r + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * r + q] - A[p * r + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < r; ++k) { const double akp = A[k * r + p], akq = A[k * r + q]; A[k * r + p] = c * akp - sn * akq; A[k * r + q] = sn * akp + c * akq; }
                for (int k = 0; k < r; ++k) { const double apk = A[p * r + k], aqk = A[q * r + k]; A[p * r + k] = c * apk - sn * aqk; A[q * r + k] = sn * apk + c * aqk; }
                for (int k = 0; k < r; ++k) { const double vkp = V[k * r + p], vkq = V[k * r + q]; V[k * r + p] = c * vkp - sn * vkq; V[k * r + q] = sn * vkp + c * vkq; }
            }
    }
    int order[kMaxR];
    for (int j = 0; j < r; ++j) order[j] = j;
    for (int a = 0; a < r; ++a) for (int b = a + 1; b < r; ++b) if (A[order[b] * r + order[b]] > A[order[a] * r + order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
    for (int j = 0; j < r; ++j) eig[j] = A[order[j] * r + order[j]];
    for (int k = 0; k < 3; ++k) for (int j = 0; j < r; ++j) W3[j * 3 + k] = V[j * r + order[k]];
}

// ---- per camera: B = (s_i R_i W3)^T, scale, unit-scale block.  blk[i][k][a] (k = row of the 3 x 3N result)
__global__ void blocks_kernel(const double* __restrict__ R, const double* __restrict__ s, const double* __restrict__ W3, int N, int r,
                              double* __restrict__ blk, double* __restrict__ s_real) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int n3 = 3 * N;
    double B[3][3] = {};
    if (W3) {
        for (int j = 0; j < r; ++j) {
            const double w0 = W3[j * 3], w1 = W3[j * 3 + 1], w2 = W3[j * 3 + 2];
            for (int a = 0; a < 3; ++a) {
                const double x = s[i] * R[(size_t)j * n3 + 3 * i + a];
                B[0][a] = fma(x, w0, B[0][a]); B[1][a] = fma(x, w1, B[1][a]); B[2][a] = fma(x, w2, B[2][a]);
            }
        }
    } else {                                               // r == 3: sR_real = sR^T exactly (:32-37)
        for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) B[k][a] = s[i] * R[(size_t)k * n3 + 3 * i + a];
    }
    double f = 0.0;
    for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) f += B[k][a] * B[k][a];
    const double sr = sqrt(f) / sqrt(3.0);
    s_real[i] = sr;
    for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) blk[(size_t)i * 9 + k * 3 + a] = B[k][a] / sr;
}

// polar factor U V^T of a 3 x 3 matrix by one-sided (Hestenes) Jacobi: rotate column pairs of A until orthogonal;
// then A = U S (columns), and U V^T = sum_k u_k v_k^T.  Returns det(U V^T) sign through *neg.
__device__ void polar3(double (&M)[3][3], double (&P)[3][3], int* neg) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) A[a][b] = M[a][b];
    for (int sweep = 0; sweep < 40; ++sweep) {
        double worst = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double al = A[0][p] * A[0][p] + A[1][p] * A[1][p] + A[2][p] * A[2][p];
                const double be = A[0][q] * A[0][q] + A[1][q] * A[1][q] + A[2][q] * A[2][q];
                const double ga = A[0][p] * A[0][q] + A[1][p] * A[1][q] + A[2][p] * A[2][q];
                if (ga == 0.0) continue;
                worst = fmax(worst, fabs(ga) / sqrt(al * be + 1e-300));
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int k = 0; k < 3; ++k) {
                    const double ap = A[k][p], aq = A[k][q];
                    A[k][p] = c * ap - sn * aq; A[k][q] = sn * ap + c * aq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - sn * vq; V[k][q] = sn * vp + c * vq;
                }
            }
        if (worst < 1e-16) break;
    }
    double U[3][3], sig[3];
    for (int k = 0; k < 3; ++k) {
        sig[k] = sqrt(A[0][k] * A[0][k] + A[1][k] * A[1][k] + A[2][k] * A[2][k]);
        const double inv = sig[k] > 0.0 ? 1.0 / sig[k] : 0.0;
        for (int a = 0; a < 3; ++a) U[a][k] = A[a][k] * inv;
    }
    for (int k = 0; k < 3; ++k) {                          // rank-deficient block: complete U by a cross product (any completion is a valid SVD)
        if (sig[k] > 1e-300) continue;
        const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        U[0][k] = U[1][k1] * U[2][k2] - U[2][k1] * U[1][k2];
        U[1][k] = U[2][k1] * U[0][k2] - U[0][k1] * U[2][k2];
        U[2][k] = U[0][k1] * U[1][k2] - U[1][k1] * U[0][k2];
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) P[a][b] = U[a][0] * V[b][0] + U[a][1] * V[b][1] + U[a][2] * V[b][2];
    const double det = P[0][0] * (P[1][1] * P[2][2] - P[1][2] * P[2][1]) - P[0][1] * (P[1][0] * P[2][2] - P[1][2] * P[2][0]) +
                       P[0][2] * (P[1][0] * P[2][1] - P[1][1] * P[2][0]);
    *neg = det < 0.0 ? 1 : 0;
}

// ---- anchoring by camera 0 (:47-48) + O(3) projection (:50-73); counts det < 0
__global__ void polar_kernel(const double* __restrict__ blk, int N, double* __restrict__ pol, int* __restrict__ negative) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int neg = 0;
    if (i < N) {
        double M[3][3], P[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {                   // M = R_1^T R_i
                double m = 0.0;
                for (int k = 0; k < 3; ++k) m = fma(blk[k * 3 + a], blk[(size_t)i * 9 + k * 3 + b], m);
                M[a][b] = m;
            }
        polar3(M, P, &neg);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) pol[(size_t)i * 9 + a * 3 + b] = P[a][b];
    }
    const unsigned m = __ballot_sync(0xffffffffu, neg);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(negative, __popc(m));       // integer count: order-independent
}

// ---- global sign (:62-63), outputs in the reference's layouts: R_out 3 x 3N column-major, X = (s R)^T as 3N x 3 column-major
__global__ void finish_kernel(const double* __restrict__ pol, const double* __restrict__ s_real, const int* __restrict__ negative, int N,
                              double* __restrict__ R_out, double* __restrict__ X) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double sign = (2 * *negative > N) ? -1.0 : 1.0;
    const int n3 = 3 * N;
    for (int k = 0; k < 3; ++k)
        for (int a = 0; a < 3; ++a) {
            const double v = sign * pol[(size_t)i * 9 + k * 3 + a];
            R_out[(size_t)(3 * i + a) * 3 + k] = v;
            X[(size_t)k * n3 + 3 * i + a] = s_real[i] * v;
        }
}

// ---- y[k][row] partial sums: Abar (rows x 3N, column-major) times X (3N x 3); columns [c0, c1) per blockIdx.y
__global__ void abar_partial_kernel(const double* __restrict__ Abar, long long rows, int n3, const double* __restrict__ X,
                                    int cols_per_split, double* __restrict__ part) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * cols_per_split, c1 = min(n3, c0 + cols_per_split);
    if (row >= rows) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
        const double v = __ldcs(Abar + (size_t)c * rows + row);             // streamed once
        a0 = fma(v, X[c], a0); a1 = fma(v, X[n3 + c], a1); a2 = fma(v, X[2 * n3 + c], a2);
    }
    double* p = part + ((size_t)blockIdx.y * rows + row) * 3;
    p[0] = a0; p[1] = a1; p[2] = a2;
}
// y_out: 3 x (rows + 1) column-major, first column zero (:79)
__global__ void abar_reduce_kernel(const double* __restrict__ part, long long rows, int nsplit, double* __restrict__ y) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row == 0) { y[0] = y[1] = y[2] = 0.0; }
    if (row >= rows) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) {
        const double* p = part + ((size_t)sp * rows + row) * 3;
        a0 += p[0]; a1 += p[1]; a2 += p[2];
    }
    double* o = y + (size_t)(row + 1) * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
}

// ---- XM^2 support: weighted squared residual of every observation (3_test_colmap_glomap.py:304-316)
//      err[o] = w[o] * || p[:, lm[o]] - ( s[cam[o]] * R_cam[o] * pt[o] + t[:, cam[o]] ) ||^2 ,  R_cam = R_real[:, 3 cam .. 3 cam + 2]
__global__ void residuals_kernel(long long n_obs, const int* __restrict__ cam, const int* __restrict__ lm, const double* __restrict__ pts,
                                 const double* __restrict__ w, const double* __restrict__ R, const double* __restrict__ s,
                                 const double* __restrict__ t, const double* __restrict__ p, double* __restrict__ err) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_obs) return;
    const int i = cam[o], k = lm[o];
    const double x0 = pts[3 * o], x1 = pts[3 * o + 1], x2 = pts[3 * o + 2];
    const double* Ri = R + (size_t)9 * i;                    // column-major 3 x 3 block: Ri[row + 3 col]
    const double si = s[i];
    double e = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double y = si * (Ri[a] * x0 + Ri[a + 3] * x1 + Ri[a + 6] * x2) + t[(size_t)3 * i + a];
        const double dlt = p[(size_t)3 * k + a] - y;
        e = fma(dlt, dlt, e);
    }
    err[o] = w[o] * e;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 8)); }
    template <class T> T* as() { return (T*)p; }
};

}  // namespace

extern "C" int xm_recover(xm_handle* h, int N, int r, const double* R, const double* s, const double* Abar, int64_t abar_rows,
                          double* R_out, double* s_out, double* y_out, double* eig_out, int* negative_out) {
    if (!h || !R || !s || !R_out || !s_out || N <= 0 || r < 3 || r > XM_MAX_RANK || (Abar && (!y_out || abar_rows <= 0))) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int n3 = 3 * N;
    const int nparts = std::min(h->num_sm, (n3 + 255) / 256);
    DevBuf dR, ds, dpart, dW3, deig, dblk, dsreal, dpol, dneg, dRout, dX, dA, dyp, dy;
    XM_CUDA(h, dR.alloc(sizeof(double) * n3 * r)); XM_CUDA(h, ds.alloc(sizeof(double) * N));
    XM_CUDA(h, dpart.alloc(sizeof(double) * nparts * r * r)); XM_CUDA(h, dW3.alloc(sizeof(double) * r * 3)); XM_CUDA(h, deig.alloc(sizeof(double) * r));
    XM_CUDA(h, dblk.alloc(sizeof(double) * 9 * N)); XM_CUDA(h, dsreal.alloc(sizeof(double) * N)); XM_CUDA(h, dpol.alloc(sizeof(double) * 9 * N));
    XM_CUDA(h, dneg.alloc(sizeof(int))); XM_CUDA(h, dRout.alloc(sizeof(double) * 9 * N)); XM_CUDA(h, dX.alloc(sizeof(double) * 9 * N));
    XM_CUDA(h, cudaMemcpyAsync(dR.p, R, sizeof(double) * n3 * r, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(ds.p, s, sizeof(double) * N, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemsetAsync(dneg.p, 0, sizeof(int), st));
    const int TB = 128, GB = (N + TB - 1) / TB;
    if (r > 3) {
        gram_partial_kernel<<<nparts, 416, sizeof(double) * 64 * r, st>>>(dR.as<double>(), ds.as<double>(), n3, r, dpart.as<double>());
        gram_eig_kernel<<<1, 128, 0, st>>>(dpart.as<double>(), nparts, r, dW3.as<double>(), deig.as<double>());
        h->launches += 2;
    }
    blocks_kernel<<<GB, TB, 0, st>>>(dR.as<double>(), ds.as<double>(), r > 3 ? dW3.as<double>() : nullptr, N, r, dblk.as<double>(), dsreal.as<double>());
    polar_kernel<<<GB, TB, 0, st>>>(dblk.as<double>(), N, dpol.as<double>(), dneg.as<int>());
    finish_kernel<<<GB, TB, 0, st>>>(dpol.as<double>(), dsreal.as<double>(), dneg.as<int>(), N, dRout.as<double>(), dX.as<double>());
    h->launches += 3;
    XM_CUDA(h, cudaGetLastError());
    if (Abar) {
        const long long rows = abar_rows;
        XM_CUDA(h, dA.alloc(sizeof(double) * (size_t)rows * n3));
        XM_CUDA(h, cudaMemcpyAsync(dA.p, Abar, sizeof(double) * (size_t)rows * n3, cudaMemcpyHostToDevice, st));
        const int rb = (int)((rows + 255) / 256);
        int nsplit = std::max(1, std::min((2 * h->num_sm + rb - 1) / rb, (n3 + 63) / 64));     // enough CTAs to fill the GPU
        const int cps = (n3 + nsplit - 1) / nsplit;
        nsplit = (n3 + cps - 1) / cps;
        XM_CUDA(h, dyp.alloc(sizeof(double) * 3 * (size_t)rows * nsplit)); XM_CUDA(h, dy.alloc(sizeof(double) * 3 * (size_t)(rows + 1)));
        abar_partial_kernel<<<dim3(rb, nsplit), 256, 0, st>>>(dA.as<double>(), rows, n3, dX.as<double>(), cps, dyp.as<double>());
        abar_reduce_kernel<<<(int)((rows + 1 + 255) / 256), 256, 0, st>>>(dyp.as<double>(), rows, nsplit, dy.as<double>());
        h->launches += 2;
        XM_CUDA(h, cudaGetLastError());
        XM_CUDA(h, cudaMemcpyAsync(y_out, dy.p, sizeof(double) * 3 * (size_t)(rows + 1), cudaMemcpyDeviceToHost, st));
    }
    XM_CUDA(h, cudaMemcpyAsync(R_out, dRout.p, sizeof(double) * 9 * N, cudaMemcpyDeviceToHost, st));
    XM_CUDA(h, cudaMemcpyAsync(s_out, dsreal.p, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    if (eig_out) {
        if (r > 3) XM_CUDA(h, cudaMemcpyAsync(eig_out, deig.p, sizeof(double) * r, cudaMemcpyDeviceToHost, st));
        else for (int j = 0; j < r; ++j) eig_out[j] = 0.0;
    }
    int neg = 0;
    XM_CUDA(h, cudaMemcpyAsync(&neg, dneg.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    XM_CUDA(h, cudaStreamSynchronize(st));
    if (negative_out) *negative_out = neg;
    return XM_OK;
}

extern "C" int xm_residuals(xm_handle* h, int64_t n_obs, int N, int M, const int* cam, const int* lm, const double* pts, const double* w,
                            const double* R_real, const double* s_real, const double* t, const double* p, double* err_out) {
    if (!h || n_obs <= 0 || N <= 0 || M <= 0 || !cam || !lm || !pts || !w || !R_real || !s_real || !t || !p || !err_out) return XM_EINVAL;
    for (int64_t o = 0; o < n_obs; ++o)
        if (cam[o] < 0 || cam[o] >= N || lm[o] < 0 || lm[o] >= M) { h->err = "xm_residuals: observation index out of range"; return XM_EINVAL; }
    XM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    DevBuf dcam, dlm, dpts, dw, dR, ds, dt, dp, derr;
    XM_CUDA(h, dcam.alloc(sizeof(int) * n_obs)); XM_CUDA(h, dlm.alloc(sizeof(int) * n_obs)); XM_CUDA(h, dpts.alloc(sizeof(double) * 3 * n_obs));
    XM_CUDA(h, dw.alloc(sizeof(double) * n_obs)); XM_CUDA(h, derr.alloc(sizeof(double) * n_obs));
    XM_CUDA(h, dR.alloc(sizeof(double) * 9 * N)); XM_CUDA(h, ds.alloc(sizeof(double) * N)); XM_CUDA(h, dt.alloc(sizeof(double) * 3 * N));
    XM_CUDA(h, dp.alloc(sizeof(double) * 3 * M));
    XM_CUDA(h, cudaMemcpyAsync(dcam.p, cam, sizeof(int) * n_obs, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(dlm.p, lm, sizeof(int) * n_obs, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(dpts.p, pts, sizeof(double) * 3 * n_obs, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(dw.p, w, sizeof(double) * n_obs, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(dR.p, R_real, sizeof(double) * 9 * N, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(ds.p, s_real, sizeof(double) * N, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(dt.p, t, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, st));
    XM_CUDA(h, cudaMemcpyAsync(dp.p, p, sizeof(double) * 3 * M, cudaMemcpyHostToDevice, st));
    residuals_kernel<<<(unsigned)((n_obs + 255) / 256), 256, 0, st>>>(n_obs, dcam.as<int>(), dlm.as<int>(), dpts.as<double>(), dw.as<double>(),
                                                                     dR.as<double>(), ds.as<double>(), dt.as<double>(), dp.as<double>(), derr.as<double>());
    XM_CUDA(h, cudaGetLastError());
    h->launches++;
    XM_CUDA(h, cudaMemcpyAsync(err_out, derr.p, sizeof(double) * n_obs, cudaMemcpyDeviceToHost, st));
    XM_CUDA(h, cudaStreamSynchronize(st));
    return XM_OK;
}
