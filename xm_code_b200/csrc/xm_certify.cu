// xm_certify.cu — optimality certificate behind xm_certify / xm_certify_ex (replaces checkeig, XM/include/XM/checkeig.h:42-368).
//
// Same mathematics as the reference, different mechanics — everything below runs on the device:
//   * Z sR comes from the library's own Q.Y kernel (+ the lambda diagonal, checkeig.h:31-40,179-182);
//   * the least-squares multipliers are solved in CLOSED FORM per camera (one thread per camera): constraint j only touches
//     its own camera's 3 rows, so the normal equations of checkeig.h:190-220 (Eigen LSCG on a 3Nr x (5N+1) sparse matrix on
//     the host) are block diagonal — one 6x6 system for camera 0 and one 5x5 per other camera;
//   * the dual slack S = Z - sum_j y_j A_j differs from Q only in its 3x3 diagonal blocks (checkeig.h:263-300), so
//     S X = Q X + blockdiag(L) X: the only O(N^2) work is the library's Q.Y product;
//   * lambda_min(S) and its eigenvector:
//       XM_CERT_DENSE      cusolverDnDsyevd on the assembled S like the reference (checkeig.h:303-318): O(N^3), one GPU, dense Q;
//       XM_CERT_ITERATIVE  block Davidson on the operator S with up to 20 columns per product (a product with 20 columns moves
//                          the same HBM bytes as one with 1: the operator is bandwidth-bound), block-Jacobi preconditioner from
//                          the 3x3 diagonal blocks of S, thick restart; the start block contains the columns of sR (exact null
//                          vectors of S at a critical point), so a tight solution is confirmed in ~10-20 products and a
//                          non-tight one yields its escape direction in ~10-30.  Works on block-CSR Q and on a communicator
//                          (xm_qy is collective and returns the full product on every rank: every rank runs the same
//                          iteration on identical numbers and takes the same decision).
//     A converged Ritz value is an upper bound of lambda_min with a small residual; like every iterative certificate (SE-Sync's
//     included) it cannot PROVE a lower bound.  XM_CERT_AUTO = dense for one-GPU dense Q with 3N <= 6000, iterative otherwise.
#include "xm_host.h"
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace {

// ------------------------------------------------------------------------------------------------ small device helpers
// S (n3 x n3, column-major, unpadded) <- Qp (padded row-major).  For symmetric Q both are the same matrix; only the
// lower triangle is referenced by syevd.
__global__ void unpack_q_kernel(const double* __restrict__ Qp, int ldq, int n3, double* __restrict__ S) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (k < n3) S[(size_t)i * n3 + k] = Qp[(size_t)i * ldq + k];     // S[k, i] = Q[i, k]
}
// S[3i+a, 3i+b] += L[i][a][b]
__global__ void add_blockdiag_kernel(double* S, int n3, const double* L, int N) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N * 9) {
        const int i = t / 9, a = (t % 9) / 3, b = t % 3;
        S[(size_t)(3 * i + b) * n3 + 3 * i + a] += L[t];
    }
}
// sR[row, j] = R[row, j] * s[row / 3]
__global__ void form_sr_kernel(const double* __restrict__ R, const double* __restrict__ s, int n3, int r, double* __restrict__ sR) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)n3 * r) { const int row = (int)(t % n3); sR[t] = R[t] * s[row / 3]; }
}

// solve the m x m SPD-ish system M y = g by Gaussian elimination with complete pivoting and a rank threshold
// (rank-deficient directions get y = 0: the basic solution of the least-squares problem)
__device__ void solve_small(int m, double* M, double* g, double* y) {
    int perm[6];
    for (int i = 0; i < m; ++i) perm[i] = i;
    double scale = 0;
    for (int i = 0; i < m * m; ++i) scale = fmax(scale, fabs(M[i]));
    const double tol = scale * 1e-13;
    int rank = m;
    for (int c = 0; c < m; ++c) {
        int pr = c, pc = c; double best = -1;
        for (int i = c; i < m; ++i) for (int j = c; j < m; ++j) if (fabs(M[i * m + j]) > best) { best = fabs(M[i * m + j]); pr = i; pc = j; }
        if (best <= tol) { rank = c; break; }
        if (pr != c) { for (int j = 0; j < m; ++j) { const double t = M[pr * m + j]; M[pr * m + j] = M[c * m + j]; M[c * m + j] = t; } const double t = g[pr]; g[pr] = g[c]; g[c] = t; }
        if (pc != c) { for (int i = 0; i < m; ++i) { const double t = M[i * m + pc]; M[i * m + pc] = M[i * m + c]; M[i * m + c] = t; } const int t = perm[pc]; perm[pc] = perm[c]; perm[c] = t; }
        for (int i = c + 1; i < m; ++i) {
            const double f = M[i * m + c] / M[c * m + c];
            for (int j = c; j < m; ++j) M[i * m + j] -= f * M[c * m + j];
            g[i] -= f * g[c];
        }
    }
    double z[6] = {0, 0, 0, 0, 0, 0};
    for (int c = rank - 1; c >= 0; --c) {
        double t = g[c];
        for (int j = c + 1; j < rank; ++j) t -= M[c * m + j] * z[j];
        z[c] = t / M[c * m + c];
    }
    for (int c = 0; c < m; ++c) y[perm[c]] = z[c];
}

// constraint bases (checkeig.h:71-161): image row a of constraint q = sum_b coef[q][a][b] * x_b
__device__ void constraint_coef(bool first, double (&coef)[6][3][3]) {
    for (int q = 0; q < 6; ++q) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) coef[q][a][b] = 0.0;
    if (first) {
        coef[0][0][0] = 1.0;                                   // E_00
        coef[1][0][1] = 0.5; coef[1][1][0] = 0.5;              // 1/2 (E_01 + E_10)
        coef[2][0][2] = 0.5; coef[2][2][0] = 0.5;              // 1/2 (E_02 + E_20)
        coef[3][1][1] = 1.0;                                   // E_11
        coef[4][1][2] = 0.5; coef[4][2][1] = 0.5;              // 1/2 (E_12 + E_21)
        coef[5][2][2] = 1.0;                                   // E_22
    } else {
        coef[0][0][0] = 0.5; coef[0][1][1] = -0.5;             // 1/2 (E_aa - E_bb)
        coef[1][1][1] = 0.5; coef[1][2][2] = -0.5;             // 1/2 (E_bb - E_cc)
        coef[2][0][1] = 0.5; coef[2][1][0] = 0.5;              // 1/2 (E_ab + E_ba)
        coef[3][0][2] = 0.5; coef[3][2][0] = 0.5;              // 1/2 (E_ac + E_ca)
        coef[4][1][2] = 0.5; coef[4][2][1] = 0.5;              // 1/2 (E_bc + E_cb)
    }
}

// One thread per camera.  In: sR and QsR = Q sR (3N x r column-major).  Out: L[i] = the 3x3 block to ADD to Q's diagonal block
// (lambda term minus sum_j y_j A_j), dterm[i] = lam (1 - x_ii^2) (checkeig.h:330-332), y0dual = y[0] + y[3] + y[5] of camera 0.
__global__ void multipliers_kernel(const double* __restrict__ sR, const double* __restrict__ QsR, int N, int n3, int r, double lam,
                                   double* __restrict__ L, double* __restrict__ dterm, double* __restrict__ y0dual) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int m = (i == 0) ? 6 : 5;
    double coef[6][3][3];
    constraint_coef(i == 0, coef);
    double xii = 0;
    for (int j = 0; j < r; ++j) { const double x = sR[(size_t)j * n3 + 3 * i]; xii += x * x; }   // ConstructZmatrixKernal: Z[3i,3i] += 2 lam (|sR_3i|^2 - 1)
    const double zc = 2.0 * lam * (xii - 1.0);
    double M[36], g[6], y[6] = {0, 0, 0, 0, 0, 0};
    for (int p = 0; p < m; ++p) { g[p] = 0; for (int q = 0; q < m; ++q) M[p * m + q] = 0; }
    for (int j = 0; j < r; ++j) {
        double x[3], w[3];
        for (int a = 0; a < 3; ++a) { x[a] = sR[(size_t)j * n3 + 3 * i + a]; w[a] = QsR[(size_t)j * n3 + 3 * i + a]; }
        w[0] += zc * x[0];
        double C[6][3];
        for (int q = 0; q < m; ++q)
            for (int a = 0; a < 3; ++a) C[q][a] = coef[q][a][0] * x[0] + coef[q][a][1] * x[1] + coef[q][a][2] * x[2];
        for (int p = 0; p < m; ++p) {
            for (int q = 0; q < m; ++q) M[p * m + q] += C[p][0] * C[q][0] + C[p][1] * C[q][1] + C[p][2] * C[q][2];
            g[p] += C[p][0] * w[0] + C[p][1] * w[1] + C[p][2] * w[2];
        }
    }
    solve_small(m, M, g, y);
    double Li[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < m; ++q)
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Li[a * 3 + b] -= y[q] * coef[q][a][b];   // S = Z - sum y A
    Li[0] += zc;
    for (int t = 0; t < 9; ++t) L[(size_t)i * 9 + t] = Li[t];
    dterm[i] = (1.0 - xii * xii) * lam;
    if (i == 0) y0dual[0] = y[0] + y[3] + y[5];                  // checkeig.h:322
}

// Y[:, c] += blockdiag(L) X[:, c]   (n3 x k column-major, ld = n3)
__global__ void apply_blockdiag_kernel(const double* __restrict__ L, const double* __restrict__ X, int N, int n3, int k, double* __restrict__ Y) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * k) return;
    const int i = (int)(t % N), c = (int)(t / N);
    const double* l = L + (size_t)i * 9;
    const double* x = X + (size_t)c * n3 + 3 * i;
    double* y = Y + (size_t)c * n3 + 3 * i;
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    y[0] += l[0] * x0 + l[1] * x1 + l[2] * x2;
    y[1] += l[3] * x0 + l[4] * x1 + l[5] * x2;
    y[2] += l[6] * x0 + l[7] * x1 + l[8] * x2;
}

// symmetric 3x3 eigendecomposition by cyclic Jacobi (A = V diag(w) V^T)
__device__ void eig3(double A[3][3], double w[3], double V[3][3]) {
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) V[a][b] = (a == b);
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
    for (int a = 0; a < 3; ++a) w[a] = A[a][a];
}
// block-Jacobi preconditioner: Dinv[i] = |Qdiag_i + L_i|^{-1} (eigenvalues by modulus, clamped: symmetric positive definite
// whatever the inertia of S — a legitimate preconditioner for the Davidson expansion also when S is indefinite)
__global__ void precond_build_kernel(const double* __restrict__ Qd, const double* __restrict__ L, int N, double* __restrict__ Dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double A[3][3], w[3], V[3][3];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) A[a][b] = 0.5 * (Qd[(size_t)i * 9 + a * 3 + b] + Qd[(size_t)i * 9 + b * 3 + a]) + 0.5 * (L[(size_t)i * 9 + a * 3 + b] + L[(size_t)i * 9 + b * 3 + a]);
    eig3(A, w, V);
    const double wm = fmax(fabs(w[0]), fmax(fabs(w[1]), fabs(w[2])));
    for (int a = 0; a < 3; ++a) w[a] = 1.0 / fmax(fabs(w[a]), fmax(1e-8 * wm, 1e-300));
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
        Dinv[(size_t)i * 9 + a * 3 + b] = V[a][0] * w[0] * V[b][0] + V[a][1] * w[1] * V[b][1] + V[a][2] * w[2] * V[b][2];
}
// W[:, c] = blockdiag(Dinv) (SX[:, c] - theta[c] X[:, c])   and   Rn2[c] += |residual|^2 (one atomic per block: used for the stop test only)
__global__ void residual_precond_kernel(const double* __restrict__ X, const double* __restrict__ SX, const double* __restrict__ theta,
                                        const double* __restrict__ Dinv, int N, int n3, int k, double* __restrict__ W) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * k) return;
    const int i = (int)(t % N), c = (int)(t / N);
    const size_t o = (size_t)c * n3 + 3 * i;
    const double th = theta[c];
    const double r0 = SX[o] - th * X[o], r1 = SX[o + 1] - th * X[o + 1], r2 = SX[o + 2] - th * X[o + 2];
    const double* d = Dinv + (size_t)i * 9;
    W[o] = d[0] * r0 + d[1] * r1 + d[2] * r2;
    W[o + 1] = d[3] * r0 + d[4] * r1 + d[5] * r2;
    W[o + 2] = d[6] * r0 + d[7] * r1 + d[8] * r2;
}
// out[c] = || A[:, c] - theta[c] B[:, c] ||_2  (theta == nullptr: || A[:, c] ||); one block per column, fixed-order tree: deterministic
__global__ void col_norm_kernel(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ theta, int n3, double* __restrict__ out) {
    __shared__ double red[256];
    const int c = blockIdx.x;
    const double th = theta ? theta[c] : 0.0;
    double acc = 0;
    for (int i = threadIdx.x; i < n3; i += blockDim.x) {
        const double v = A[(size_t)c * n3 + i] - (theta ? th * B[(size_t)c * n3 + i] : 0.0);
        acc += v * v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) out[c] = sqrt(red[0]);
}
// A[:, c] *= f[c]
__global__ void col_scale_kernel(double* __restrict__ A, const double* __restrict__ f, int n3, int k) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)n3 * k) A[t] *= f[t / n3];
}
__global__ void symmetrize_kernel(double* H, int m, int ld) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m * m) { const int i = t % m, j = t / m; if (i < j) { const double v = 0.5 * (H[(size_t)j * ld + i] + H[(size_t)i * ld + j]); H[(size_t)j * ld + i] = v; H[(size_t)i * ld + j] = v; } }
}

// stream-ordered allocations: cudaFree would synchronise the whole device, and on a loop-back communicator (two members on one
// GPU) a peer's persistent kernel may already be waiting for this member's next collective call
struct DevBuf {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    ~DevBuf() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    template <class T> T* get(size_t n) { void* p = nullptr; if (cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st) != cudaSuccess) { cudaGetLastError(); return nullptr; } ptrs.push_back(p); return (T*)p; }
};

#define CERT_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); cudaGetLastError(); return XM_ECUDA; } } while (0)
#define CERT_BLAS(call) do { if ((call) != CUBLAS_STATUS_SUCCESS) { h->err = #call; return XM_ECUDA; } } while (0)
#define CERT_SOLV(call) do { if ((call) != CUSOLVER_STATUS_SUCCESS) { h->err = #call; return XM_ECUDA; } } while (0)

inline unsigned nblk(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// S X for a block of k >= 1 columns (device, ld = n3): slices of 3..20 columns through the product kernel + the block diagonal
int apply_S(xm_handle* h, const double* L, const double* X, double* Y, int k, double* pad3_in, double* pad3_out, int* products) {
    const int n3 = h->n3, N = h->N;
    const int wmax = (h->world > 1) ? std::max(3, std::min(h->comm_maxr, XM_MAX_RANK)) : XM_MAX_RANK;    // a communicator was sized for comm_maxr columns
    if (k < 3) {            // the kernel's minimum width: pad with zero columns
        CERT_CUDA(cudaMemsetAsync(pad3_in, 0, (size_t)n3 * 3 * sizeof(double), h->stream));
        CERT_CUDA(cudaMemcpyAsync(pad3_in, X, (size_t)n3 * k * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        int rc = xm_qy_dev(h, 3, 1.0, pad3_in, pad3_out);
        if (rc) return rc;
        CERT_CUDA(cudaMemcpyAsync(Y, pad3_out, (size_t)n3 * k * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        ++*products;
    } else {
        int j0 = 0;
        while (j0 < k) {
            int take = std::min(wmax, k - j0);
            const int rem = k - j0 - take;
            if (rem > 0 && rem < 3) take -= (3 - rem);
            int rc = xm_qy_dev(h, take, 1.0, X + (size_t)j0 * n3, Y + (size_t)j0 * n3);
            if (rc) return rc;
            ++*products;
            j0 += take;
        }
    }
    apply_blockdiag_kernel<<<nblk((long long)N * k), 256, 0, h->stream>>>(L, X, N, n3, k, Y);
    CERT_CUDA(cudaGetLastError());
    return XM_OK;
}

// symmetric eigendecomposition of the m x m matrix A (device, column-major, ld): eigenvalues ascending to w_host, vectors in place
int small_eigh(xm_handle* h, cusolverDnHandle_t cs, double* A, int m, int ld, double* w_dev, double* work, int lwork, int* info, double* w_host) {
    CERT_SOLV(cusolverDnDsyevd(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, A, ld, w_dev, work, lwork, info));
    int hinfo = 0;
    CERT_CUDA(cudaMemcpyAsync(w_host, w_dev, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CERT_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CERT_CUDA(cudaStreamSynchronize(h->stream));
    if (hinfo != 0) { h->err = "small syevd did not converge"; return XM_ECUDA; }
    return XM_OK;
}

struct Blas { cublasHandle_t cb = nullptr; cusolverDnHandle_t cs = nullptr; };     // views of the handle's cached library handles

// ------------------------------------------------------------------------------------------------ block Davidson
// lowest eigenpairs of S = Q + blockdiag(L).  sR (n3 x r): the near-null vectors that seed the block.  v_out_dev: n3.
int davidson_min_eig(xm_handle* h, Blas& bl, const double* L, const double* Dinv, const double* sR, int r, double* v_out_dev,
                     double* min_eig, int* products_out, int* converged_out, double* residual_out) {
    const int n3 = h->n3, N = h->N;
    const int nconv = std::min(r + 1, n3);                        // pairs that must converge: the r near-null ones and one more
    int b = std::max(8, std::min(r + 5, 24));                     // block width (<= 2 product slices)
    b = std::min(b, n3);
    const int mmax = std::min(n3, 8 * b);
    const int maxprod = 400;
    DevBuf mem(h->stream);
    const size_t col = (size_t)n3;
    double* V = mem.get<double>(col * mmax); double* SV = mem.get<double>(col * mmax);
    double* X = mem.get<double>(col * b); double* SX = mem.get<double>(col * b);
    double* W = mem.get<double>(col * std::max(b, 3)); double* SW = mem.get<double>(col * std::max(b, 3));
    double* T1 = mem.get<double>(col * std::max(2 * b, 3)); double* T2 = mem.get<double>(col * std::max(2 * b, 3));
    double* H = mem.get<double>((size_t)mmax * mmax); double* Cm = mem.get<double>((size_t)mmax * mmax);
    double* G = mem.get<double>((size_t)b * b);
    double* wdev = mem.get<double>(mmax); double* nrm = mem.get<double>(mmax); double* fac = mem.get<double>(mmax);
    int* info = mem.get<int>(1);
    if (!V || !SV || !X || !SX || !W || !SW || !T1 || !T2 || !H || !Cm || !G || !wdev || !nrm || !fac || !info) { h->err = "certificate workspace cudaMalloc failed"; return XM_ENOMEM; }
    int lwork = 0, lw2 = 0;
    CERT_SOLV(cusolverDnDsyevd_bufferSize(bl.cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, mmax, H, mmax, wdev, &lwork));
    CERT_SOLV(cusolverDnDsyevd_bufferSize(bl.cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, b, G, b, wdev, &lw2));
    lwork = std::max(lwork, lw2);
    for (int m = 1; m <= mmax; ++m) {     // syevd's workspace is not monotone in the size on every version: take the maximum
        int q = 0;
        if (cusolverDnDsyevd_bufferSize(bl.cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, H, mmax, wdev, &q) == CUSOLVER_STATUS_SUCCESS) lwork = std::max(lwork, q);
    }
    double* work = mem.get<double>((size_t)lwork);
    if (!work) { h->err = "certificate workspace cudaMalloc failed"; return XM_ENOMEM; }
    std::vector<double> wh(mmax), nh(mmax), fh(mmax);
    const double one = 1.0, zero = 0.0, mone = -1.0;

    // start block: the columns of sR, then pseudo-random columns (fixed LCG: identical on every rank and every run)
    {
        std::vector<double> x0(col * b);
        unsigned long long st = 0x9E3779B97F4A7C15ull;
        for (size_t t = 0; t < x0.size(); ++t) { st = st * 6364136223846793005ull + 1442695040888963407ull; x0[t] = ((double)(st >> 11) / 9007199254740992.0) - 0.5; }
        CERT_CUDA(cudaMemcpyAsync(W, x0.data(), x0.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CERT_CUDA(cudaMemcpyAsync(W, sR, col * std::min(r, b) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CERT_CUDA(cudaStreamSynchronize(h->stream));
    }
    int m = 0, k = b, products = 0, converged = 0;
    double theta0 = 0.0, res_max = 0.0;
    for (int iter = 0; iter < 10000; ++iter) {
        // ---- orthonormalise the k new columns in W against V[:, :m] and among themselves (two passes; SVQB with dropping)
        int kk = k;
        for (int pass = 0; pass < 2 && kk > 0; ++pass) {
            if (m > 0) {
                for (int rep = 0; rep < 2; ++rep) {
                    CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_T, CUBLAS_OP_N, m, kk, n3, &one, V, n3, W, n3, &zero, Cm, mmax));
                    CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_N, CUBLAS_OP_N, n3, kk, m, &mone, V, n3, Cm, mmax, &one, W, n3));
                }
            }
            col_norm_kernel<<<kk, 256, 0, h->stream>>>(W, nullptr, nullptr, n3, nrm);
            CERT_CUDA(cudaMemcpyAsync(nh.data(), nrm, kk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CERT_CUDA(cudaStreamSynchronize(h->stream));
            for (int c = 0; c < kk; ++c) fh[c] = (nh[c] > 1e-280 && std::isfinite(nh[c])) ? 1.0 / nh[c] : 0.0;
            CERT_CUDA(cudaMemcpyAsync(fac, fh.data(), kk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            col_scale_kernel<<<nblk((long long)n3 * kk), 256, 0, h->stream>>>(W, fac, n3, kk);
            CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_T, CUBLAS_OP_N, kk, kk, n3, &one, W, n3, W, n3, &zero, G, b));
            int rc = small_eigh(h, bl.cs, G, kk, b, wdev, work, lwork, info, wh.data());
            if (rc) return rc;
            const double wmax = wh[kk - 1];
            int j0 = 0;
            while (j0 < kk && !(wh[j0] > 1e-10 * wmax && wh[j0] > 0)) ++j0;       // ascending: the kept directions are a suffix
            const int keep = kk - j0;
            if (keep <= 0 || !(wmax > 0)) { kk = 0; break; }
            for (int c = 0; c < kk; ++c) fh[c] = (c >= j0) ? 1.0 / std::sqrt(wh[c]) : 0.0;
            CERT_CUDA(cudaMemcpyAsync(fac, fh.data(), kk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            col_scale_kernel<<<nblk((long long)b * kk), 256, 0, h->stream>>>(G, fac, b, kk);        // G is kk x kk with ld b: scale whole columns
            CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_N, CUBLAS_OP_N, n3, keep, kk, &one, W, n3, G + (size_t)j0 * b, b, &zero, T1, n3));
            CERT_CUDA(cudaMemcpyAsync(W, T1, col * keep * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            kk = keep;
        }
        if (kk == 0) break;                                       // nothing new to add: the basis cannot be improved
        // ---- one operator application on the new block, append
        int rc = apply_S(h, L, W, SW, kk, T1, T2, &products);
        if (rc) return rc;
        CERT_CUDA(cudaMemcpyAsync(V + col * m, W, col * kk * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CERT_CUDA(cudaMemcpyAsync(SV + col * m, SW, col * kk * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        m += kk;
        // ---- Rayleigh-Ritz on the whole basis
        CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_T, CUBLAS_OP_N, m, m, n3, &one, V, n3, SV, n3, &zero, H, mmax));
        symmetrize_kernel<<<nblk((long long)m * m), 256, 0, h->stream>>>(H, m, mmax);
        rc = small_eigh(h, bl.cs, H, m, mmax, wdev, work, lwork, info, wh.data());
        if (rc) return rc;
        const int kb = std::min(b, m);
        CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_N, CUBLAS_OP_N, n3, kb, m, &one, V, n3, H, mmax, &zero, X, n3));
        CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_N, CUBLAS_OP_N, n3, kb, m, &one, SV, n3, H, mmax, &zero, SX, n3));
        col_norm_kernel<<<kb, 256, 0, h->stream>>>(SX, X, wdev, n3, nrm);
        CERT_CUDA(cudaMemcpyAsync(nh.data(), nrm, kb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CERT_CUDA(cudaStreamSynchronize(h->stream));
        const double scale = std::max(std::fabs(wh[0]), std::fabs(wh[m - 1]));
        const double tol = std::max(1e-7 * scale, 1e-9);
        theta0 = wh[0]; res_max = 0.0;
        for (int c = 0; c < std::min(nconv, kb); ++c) res_max = std::max(res_max, nh[c]);
        if (kb >= nconv && res_max < tol) { converged = 1; break; }
        if (products >= maxprod) break;
        // ---- thick restart: keep the 2b lowest Ritz vectors
        if (m + kb > mmax) {
            const int k2 = std::min(2 * b, m);
            CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_N, CUBLAS_OP_N, n3, k2, m, &one, V, n3, H, mmax, &zero, T1, n3));
            CERT_BLAS(cublasDgemm(bl.cb, CUBLAS_OP_N, CUBLAS_OP_N, n3, k2, m, &one, SV, n3, H, mmax, &zero, T2, n3));
            CERT_CUDA(cudaMemcpyAsync(V, T1, col * k2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            CERT_CUDA(cudaMemcpyAsync(SV, T2, col * k2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            m = k2;
        }
        // ---- expansion block: preconditioned residuals of the kb lowest Ritz pairs
        residual_precond_kernel<<<nblk((long long)N * kb), 256, 0, h->stream>>>(X, SX, wdev, Dinv, N, n3, kb, W);
        CERT_CUDA(cudaGetLastError());
        k = kb;
    }
    CERT_CUDA(cudaMemcpyAsync(v_out_dev, X, col * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CERT_CUDA(cudaStreamSynchronize(h->stream));
    *min_eig = theta0; *products_out = products; *converged_out = converged; *residual_out = res_max;
    return XM_OK;
}

}  // namespace

extern "C" int xm_op_diag_blocks_dev(xm_handle* h, double* out9N_dev);      // xm_capi.cu
int xm_internal_libs(xm_handle* h, void** cublas_out, void** cusolver_out);    // xm_capi.cu: the handle's cached cuBLAS / cuSOLVER handles

extern "C" int xm_certify_ex(xm_handle* h, int r, const double* R, const double* s, double lam, double primal, int method,
                             double* v_out, xm_cert_info* out) {
    XmRange nvtx_range("xm_certify");
    if (!h || !R || !s) return XM_EINVAL;
    if (r < 3 || r > XM_MAX_RANK) return XM_EINVAL;
    if (h->N <= 0 || (!h->is_bsr && !h->Qp)) { h->err = "no Q set"; return XM_EINVAL; }
    const int N = h->N, n3 = h->n3;
    if (method == XM_CERT_AUTO) method = (h->world == 1 && !h->is_bsr && n3 <= 6000) ? XM_CERT_DENSE : XM_CERT_ITERATIVE;
    if (method == XM_CERT_DENSE && (h->is_bsr || h->world > 1)) { h->err = "the dense certificate needs a dense Q on one GPU (use XM_CERT_ITERATIVE)"; return XM_EUNSUPPORTED; }
    if (method != XM_CERT_DENSE && method != XM_CERT_ITERATIVE) return XM_EINVAL;
    XM_CUDA(h, cudaSetDevice(h->device));
    struct Events { cudaEvent_t a = nullptr, b = nullptr; ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } ev;     // freed on every exit path
    XM_CUDA(h, cudaEventCreate(&ev.a)); XM_CUDA(h, cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    XM_CUDA(h, cudaEventRecord(e0, h->stream));
    DevBuf mem(h->stream);
    const size_t col = (size_t)n3;
    double* dR = mem.get<double>(col * r); double* ds = mem.get<double>(N); double* sR = mem.get<double>(col * r); double* QsR = mem.get<double>(col * r);
    double* L = mem.get<double>((size_t)N * 9); double* dterm = mem.get<double>(N); double* y0 = mem.get<double>(1); double* vdev = mem.get<double>(col);
    if (!dR || !ds || !sR || !QsR || !L || !dterm || !y0 || !vdev) { h->err = "certificate cudaMalloc failed"; return XM_ENOMEM; }
    CERT_CUDA(cudaMemcpyAsync(dR, R, col * r * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CERT_CUDA(cudaMemcpyAsync(ds, s, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    form_sr_kernel<<<nblk((long long)n3 * r), 256, 0, h->stream>>>(dR, ds, n3, r, sR);
    CERT_CUDA(cudaGetLastError());
    int rc = xm_qy_dev(h, r, 1.0, sR, QsR);                       // Q sR through the product kernel (collective on a communicator)
    if (rc) return rc;
    multipliers_kernel<<<nblk(N, 128), 128, 0, h->stream>>>(sR, QsR, N, n3, r, lam, L, dterm, y0);
    CERT_CUDA(cudaGetLastError());
    std::vector<double> dth(N);
    double y0h = 0;
    CERT_CUDA(cudaMemcpyAsync(dth.data(), dterm, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CERT_CUDA(cudaMemcpyAsync(&y0h, y0, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CERT_CUDA(cudaStreamSynchronize(h->stream));
    double dual = y0h;
    for (int i = 0; i < N; ++i) dual += dth[i];                   // checkeig.h:330-332 (fixed order)
    double w0 = 0, resid = 0;
    int products = 1, converged = 1;
    Blas bl;
    int lrc = xm_internal_libs(h, (void**)&bl.cb, (void**)&bl.cs);
    if (lrc) return lrc;
    if (method == XM_CERT_DENSE) {
        // dual slack on the device + full symmetric eigendecomposition, like checkeig.h:303-318
        double* S = mem.get<double>(col * n3); double* W = mem.get<double>(n3); int* info = mem.get<int>(1);
        if (!S || !W || !info) { h->err = "dense certificate cudaMalloc failed (use XM_CERT_ITERATIVE)"; return XM_ENOMEM; }
        unpack_q_kernel<<<dim3((n3 + 255) / 256, n3), 256, 0, h->stream>>>(h->Qp, h->ldq, n3, S);
        add_blockdiag_kernel<<<nblk((long long)N * 9), 256, 0, h->stream>>>(S, n3, L, N);
        CERT_CUDA(cudaGetLastError());
        int lwork = 0;
        CERT_SOLV(cusolverDnDsyevd_bufferSize(bl.cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n3, S, n3, W, &lwork));
        double* work = mem.get<double>((size_t)std::max(lwork, 1));
        if (!work) { h->err = "syevd workspace cudaMalloc failed"; return XM_ENOMEM; }
        CERT_SOLV(cusolverDnDsyevd(bl.cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n3, S, n3, W, work, lwork, info));
        int hinfo = 0;
        CERT_CUDA(cudaMemcpyAsync(&w0, W, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CERT_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CERT_CUDA(cudaMemcpyAsync(vdev, S, col * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));      // :318
        CERT_CUDA(cudaStreamSynchronize(h->stream));
        if (hinfo != 0) { h->err = "syevd did not converge"; return XM_ECUDA; }
    } else {
        double* Qd = mem.get<double>((size_t)N * 9); double* Dinv = mem.get<double>((size_t)N * 9);
        if (!Qd || !Dinv) { h->err = "certificate cudaMalloc failed"; return XM_ENOMEM; }
        rc = xm_op_diag_blocks_dev(h, Qd);                        // 3x3 diagonal blocks of Q (collective on a communicator)
        if (rc) return rc;
        precond_build_kernel<<<nblk(N, 128), 128, 0, h->stream>>>(Qd, L, N, Dinv);
        CERT_CUDA(cudaGetLastError());
        rc = davidson_min_eig(h, bl, L, Dinv, sR, r, vdev, &w0, &products, &converged, &resid);
        if (rc) return rc;
        products += 1;
    }
    if (v_out) CERT_CUDA(cudaMemcpyAsync(v_out, vdev, col * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    XM_CUDA(h, cudaEventRecord(e1, h->stream));
    CERT_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gap = primal - dual - 3.0 * N * std::fmin(0.0, w0);      // :334-336 (bar_s = 1)
    const double bound = (N > 2000) ? 1e-3 : 1e-4;                         // :349-358 (later tiers unreachable, quirk Q5)
    const int certified = (gap / primal < 1e-3 || w0 > -bound) ? 1 : 0;    // :360
    if (out) {
        memset(out, 0, sizeof(*out));
        out->certified = certified; out->method = method; out->products = products; out->converged = converged;
        out->min_eig = w0; out->dual = dual; out->gap = gap; out->residual = resid; out->ms = ms;
    }
    if (h->opt.verbose) {
        printf("The min eig is: %1.3e \n", w0);
        printf("Primal value: %g\nDual value: %g\nOptimility gap: %g\n", primal, dual, gap);
        printf(certified ? "BM finished with rank %d\n" : "BM order plus one\n", r);
        fflush(stdout);
    }
    return XM_OK;
}

extern "C" int xm_certify(xm_handle* h, int r, const double* R, const double* s, double lam, double primal,
                          double* v_out, double* min_eig_out, double* dual_out, double* gap_out, int* certified_out) {
    xm_cert_info ci;
    const int rc = xm_certify_ex(h, r, R, s, lam, primal, XM_CERT_AUTO, v_out, &ci);
    if (rc) return rc;
    if (min_eig_out) *min_eig_out = ci.min_eig;
    if (dual_out) *dual_out = ci.dual;
    if (gap_out) *gap_out = ci.gap;
    if (certified_out) *certified_out = ci.certified;
    return XM_OK;
}
