// tools/qy_tune.cu — micro-benchmark of dense FP64 Q.Y variants on B200 (tuning tool, not shipped in the library).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/qy_tune.cu -o tools/qy_tune
// Run  : tools/qy_tune [N_cameras=1723] [r=3] [iters=50]
// Prints one line per variant: achieved algorithmic GB/s (72 N^2 + 48 N r bytes per product) and max abs error.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double2 ld_na(const double* p) {
    double2 v; asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v;
}
__device__ __forceinline__ double2 ld_plain(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 ld_cs(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct P { const double* Q; const double* Xt; double* out; int N, n3, ldq, r, G, KS; };

// ---------------------------------------------------------------------------------------------------------------
// V_cam: warp per camera (3 rows) x interleaved k-split over KS warps of the CTA; LD: 0 = nc.no_allocate, 1 = plain, 2 = cs
template <int RP, int UNROLL, int LD, int NT>
__global__ void __launch_bounds__(NT, 1) k_cam(const P p) {
    __shared__ double red[(NT / 32) * 3 * RP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = NT / 32;
    const int KS = p.KS, CB = NW / KS, cslot = warp / KS, ks = warp % KS;
    const int lo = (int)((long long)blockIdx.x * p.N / p.G), hi = (int)((long long)(blockIdx.x + 1) * p.N / p.G);
    const size_t ldq = p.ldq;
    for (int b0 = lo; b0 < hi; b0 += CB) {
        const int cam = b0 + cslot;
        double acc[3][RP];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < RP; ++j) acc[a][j] = 0.0;
        if (cam < hi) {
            const double* q0 = p.Q + (size_t)(3 * cam) * ldq;
#pragma unroll UNROLL
            for (int k = 64 * ks + 2 * lane; k < p.ldq; k += 64 * KS) {
                double2 a0, a1, a2;
                if (LD == 0) { a0 = ld_na(q0 + k); a1 = ld_na(q0 + ldq + k); a2 = ld_na(q0 + 2 * ldq + k); }
                else if (LD == 1) { a0 = ld_plain(q0 + k); a1 = ld_plain(q0 + ldq + k); a2 = ld_plain(q0 + 2 * ldq + k); }
                else { a0 = ld_cs(q0 + k); a1 = ld_cs(q0 + ldq + k); a2 = ld_cs(q0 + 2 * ldq + k); }
#pragma unroll
                for (int j = 0; j < RP; ++j) {
                    const double2 x = *reinterpret_cast<const double2*>(p.Xt + (size_t)j * ldq + k);
                    acc[0][j] = fma(a0.x, x.x, acc[0][j]); acc[0][j] = fma(a0.y, x.y, acc[0][j]);
                    acc[1][j] = fma(a1.x, x.x, acc[1][j]); acc[1][j] = fma(a1.y, x.y, acc[1][j]);
                    acc[2][j] = fma(a2.x, x.x, acc[2][j]); acc[2][j] = fma(a2.y, x.y, acc[2][j]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < RP; ++j) { double v = wsum(acc[a][j]); if (lane == 0) red[(warp * 3 + a) * RP + j] = v; }
        __syncthreads();
        const int nvalid = min(CB, hi - b0);
        for (int e = threadIdx.x; e < nvalid * 3 * RP; e += NT) {
            const int q = e / (3 * RP), a = (e / RP) % 3, j = e % RP;
            double t = 0; for (int kk = 0; kk < KS; ++kk) t += red[((q * KS + kk) * 3 + a) * RP + j];
            p.out[(size_t)(3 * (b0 + q) + a) * RP + j] = t;
        }
        __syncthreads();
    }
}

// V_row: one row per warp, rows strided over all warps of the grid
template <int RP, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_row(const P p) {
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * 512 + threadIdx.x) >> 5, nw = gridDim.x * 16;
    const size_t ldq = p.ldq;
    for (int row = gw; row < p.n3; row += nw) {
        double acc[RP];
#pragma unroll
        for (int j = 0; j < RP; ++j) acc[j] = 0;
        const double* q0 = p.Q + (size_t)row * ldq;
#pragma unroll UNROLL
        for (int k = 2 * lane; k < p.ldq; k += 64) {
            const double2 a0 = ld_na(q0 + k);
#pragma unroll
            for (int j = 0; j < RP; ++j) {
                const double2 x = *reinterpret_cast<const double2*>(p.Xt + (size_t)j * ldq + k);
                acc[j] = fma(a0.x, x.x, acc[j]); acc[j] = fma(a0.y, x.y, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < RP; ++j) { double v = wsum(acc[j]); if (lane == 0) p.out[(size_t)row * RP + j] = v; }
    }
}

// V_read: pure read-bandwidth probe over Q (contiguous, like a copy kernel's read side)
__global__ void __launch_bounds__(512, 1) k_read(const P p) {
    const size_t total = (size_t)p.n3 * p.ldq / 2;     // double2 units
    const double2* q = reinterpret_cast<const double2*>(p.Q);
    double s = 0;
    size_t i = (size_t)blockIdx.x * 512 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 512;
#pragma unroll 8
    for (; i < total; i += stride) { double2 v = ld_na(reinterpret_cast<const double*>(q + i)); s += v.x + v.y; }
    if (s == 1.2345e300) p.out[0] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// V_tma: CTA streams row tiles [ROWS x KC] + operand chunk into a shared-memory ring with cp.async.bulk (TMA 1-D bulk),
// mbarrier full/empty pairs; each consumer warp owns 3 rows (one camera) of the tile.  Spin loops are bounded.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(unsigned long long* b, unsigned parity) {
    for (unsigned it = 0; it < (1u << 22); ++it) if (mbar_try_wait(b, parity)) return true;
    return false;   // bounded: never hang the GPU
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int RP, int KC, int STAGES, int MAXCAM>
__global__ void __launch_bounds__(32 * (MAXCAM + 1), 1) k_tma(const P p, int* err) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWS = 3 * MAXCAM;
    constexpr int STAGE_DOUBLES = (ROWS + RP) * KC;
    double* ring = reinterpret_cast<double*>(smem_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)STAGES * STAGE_DOUBLES);
    unsigned long long* empty = full + STAGES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lo = (int)((long long)blockIdx.x * p.N / p.G), hi = (int)((long long)(blockIdx.x + 1) * p.N / p.G);
    const int ncam = hi - lo;                         // <= MAXCAM (host guarantees)
    const int nchunks = p.ldq / KC;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MAXCAM); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == MAXCAM) {                              // producer warp
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES; const unsigned ph = (c / STAGES) & 1;
                if (c >= STAGES && !mbar_wait(&empty[s], ph ^ 1)) { *err = 1; return; }
                double* st = ring + (size_t)s * STAGE_DOUBLES;
                mbar_expect_tx(&full[s], (unsigned)((3 * ncam + RP) * KC * sizeof(double)));
                for (int j = 0; j < RP; ++j) bulk_g2s(st + (size_t)(ROWS + j) * KC, p.Xt + (size_t)j * p.ldq + (size_t)c * KC, KC * 8, &full[s]);
                for (int rr = 0; rr < 3 * ncam; ++rr) bulk_g2s(st + (size_t)rr * KC, p.Q + (size_t)(3 * lo + rr) * p.ldq + (size_t)c * KC, KC * 8, &full[s]);
            }
        }
        return;
    }
    // consumer warps: warp w owns camera lo + w
    double acc[3][RP];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < RP; ++j) acc[a][j] = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c % STAGES; const unsigned ph = (c / STAGES) & 1;
        if (!mbar_wait(&full[s], ph)) { *err = 2; return; }
        if (warp < ncam) {
            const double* st = ring + (size_t)s * STAGE_DOUBLES;
            const double* q0 = st + (size_t)(3 * warp) * KC;
            const double* xs = st + (size_t)ROWS * KC;
#pragma unroll
            for (int k = 2 * lane; k < KC; k += 64) {
                const double2 a0 = *reinterpret_cast<const double2*>(q0 + k);
                const double2 a1 = *reinterpret_cast<const double2*>(q0 + KC + k);
                const double2 a2 = *reinterpret_cast<const double2*>(q0 + 2 * KC + k);
#pragma unroll
                for (int j = 0; j < RP; ++j) {
                    const double2 x = *reinterpret_cast<const double2*>(xs + (size_t)j * KC + k);
                    acc[0][j] = fma(a0.x, x.x, acc[0][j]); acc[0][j] = fma(a0.y, x.y, acc[0][j]);
                    acc[1][j] = fma(a1.x, x.x, acc[1][j]); acc[1][j] = fma(a1.y, x.y, acc[1][j]);
                    acc[2][j] = fma(a2.x, x.x, acc[2][j]); acc[2][j] = fma(a2.y, x.y, acc[2][j]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (warp < ncam) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < RP; ++j) { double v = wsum(acc[a][j]); if (lane == 0) p.out[(size_t)(3 * (lo + warp) + a) * RP + j] = v; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// V_tma2: like V_tma but (1) all 32 lanes of the producer warp issue bulk copies in parallel (one row segment each),
// (2) runtime chunk width KC (multiple of 64) with a short tail chunk, (3) cameras in batches of NWC per CTA.
template <int RP, int NWC>
__global__ void __launch_bounds__(32 * (NWC + 1), 1) k_tma2(const P p, const int KC, const int STAGES, int* err) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWS = 3 * NWC;
    const int stage_doubles = (ROWS + RP) * KC;
    double* ring = reinterpret_cast<double*>(smem_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)STAGES * stage_doubles);
    unsigned long long* empty = full + STAGES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lo = (int)((long long)blockIdx.x * p.N / p.G), hi = (int)((long long)(blockIdx.x + 1) * p.N / p.G);
    const int nchunks = (p.ldq + KC - 1) / KC;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == NWC) {                                 // producer warp
        int g = 0;
        for (int b0 = lo; b0 < hi; b0 += NWC) {
            const int nb = min(NWC, hi - b0);
            for (int c = 0; c < nchunks; ++c, ++g) {
                const int s = g % STAGES; const unsigned ph = (g / STAGES) & 1;
                const int kc_len = min(KC, p.ldq - c * KC);
                if (g >= STAGES) { int ok = 1; if (lane == 0) ok = mbar_wait(&empty[s], ph ^ 1); ok = __shfl_sync(0xffffffffu, ok, 0); if (!ok) { if (lane == 0) *err = 1; return; } }
                double* st = ring + (size_t)s * stage_doubles;
                if (lane == 0) mbar_expect_tx(&full[s], (unsigned)((3 * nb + RP) * kc_len * sizeof(double)));
                __syncwarp();
                for (int rr = lane; rr < 3 * nb + RP; rr += 32) {
                    const bool isq = rr < 3 * nb;
                    const double* src = isq ? p.Q + (size_t)(3 * b0 + rr) * p.ldq + (size_t)c * KC
                                            : p.Xt + (size_t)(rr - 3 * nb) * p.ldq + (size_t)c * KC;
                    double* dst = st + (size_t)(isq ? rr : ROWS + (rr - 3 * nb)) * KC;
                    bulk_g2s(dst, src, (unsigned)(kc_len * 8), &full[s]);
                }
            }
        }
        return;
    }
    int g = 0;
    for (int b0 = lo; b0 < hi; b0 += NWC) {
        const int nb = min(NWC, hi - b0);
        double acc[3][RP];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < RP; ++j) acc[a][j] = 0.0;
        for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % STAGES; const unsigned ph = (g / STAGES) & 1;
            const int kc_len = min(KC, p.ldq - c * KC);
            if (!mbar_wait(&full[s], ph)) { if (lane == 0) *err = 2; return; }
            if (warp < nb) {
                const double* st = ring + (size_t)s * stage_doubles;
                const double* q0 = st + (size_t)(3 * warp) * KC;
                const double* xs = st + (size_t)ROWS * KC;
#pragma unroll 2
                for (int k = 2 * lane; k < kc_len; k += 64) {
                    const double2 a0 = *reinterpret_cast<const double2*>(q0 + k);
                    const double2 a1 = *reinterpret_cast<const double2*>(q0 + KC + k);
                    const double2 a2 = *reinterpret_cast<const double2*>(q0 + 2 * KC + k);
#pragma unroll
                    for (int j = 0; j < RP; ++j) {
                        const double2 x = *reinterpret_cast<const double2*>(xs + (size_t)j * KC + k);
                        acc[0][j] = fma(a0.x, x.x, acc[0][j]); acc[0][j] = fma(a0.y, x.y, acc[0][j]);
                        acc[1][j] = fma(a1.x, x.x, acc[1][j]); acc[1][j] = fma(a1.y, x.y, acc[1][j]);
                        acc[2][j] = fma(a2.x, x.x, acc[2][j]); acc[2][j] = fma(a2.y, x.y, acc[2][j]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (warp < nb) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < RP; ++j) { double v = wsum(acc[a][j]); if (lane == 0) p.out[(size_t)(3 * (b0 + warp) + a) * RP + j] = v; }
        }
    }
}

// V_tma3: 2-D tensor-map TMA (cp.async.bulk.tensor.2d): one op per camera per chunk ([3 rows x KC cols] box) + one
// op for the operand chunk ([RP rows x KC]); TMA cost is per op (~85 cycles), so ops must be large.
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
template <int RP, int NWC, int KC>
__global__ void __launch_bounds__(32 * (NWC + 1), 1) k_tma3(const P p, const __grid_constant__ CUtensorMap mapQ,
                                                            const __grid_constant__ CUtensorMap mapX, const int STAGES, int* err) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWS = 3 * NWC;
    constexpr int stage_doubles = (ROWS + RP) * KC;
    double* ring = reinterpret_cast<double*>(smem_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)STAGES * stage_doubles);
    unsigned long long* empty = full + STAGES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lo = (int)((long long)blockIdx.x * p.N / p.G), hi = (int)((long long)(blockIdx.x + 1) * p.N / p.G);
    const int nchunks = (p.ldq + KC - 1) / KC;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == NWC) {
        int g = 0;
        for (int b0 = lo; b0 < hi; b0 += NWC) {
            const int nb = min(NWC, hi - b0);
            for (int c = 0; c < nchunks; ++c, ++g) {
                const int s = g % STAGES; const unsigned ph = (g / STAGES) & 1;
                if (g >= STAGES) { int ok = 1; if (lane == 0) ok = mbar_wait(&empty[s], ph ^ 1); ok = __shfl_sync(0xffffffffu, ok, 0); if (!ok) { if (lane == 0) *err = 1; return; } }
                double* st = ring + (size_t)s * stage_doubles;
                if (lane == 0) mbar_expect_tx(&full[s], (unsigned)((3 * nb + RP) * KC * sizeof(double)));   // OOB columns are zero-filled but counted
                __syncwarp();
                if (lane < nb) tma_2d(st + (size_t)(3 * lane) * KC, &mapQ, c * KC, 3 * (b0 + lane), &full[s]);
                if (lane == 31) tma_2d(st + (size_t)ROWS * KC, &mapX, c * KC, 0, &full[s]);
            }
        }
        return;
    }
    int g = 0;
    for (int b0 = lo; b0 < hi; b0 += NWC) {
        const int nb = min(NWC, hi - b0);
        double acc[3][RP];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < RP; ++j) acc[a][j] = 0.0;
        for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % STAGES; const unsigned ph = (g / STAGES) & 1;
            if (!mbar_wait(&full[s], ph)) { if (lane == 0) *err = 2; return; }
            if (warp < nb) {
                const double* st = ring + (size_t)s * stage_doubles;
                const double* q0 = st + (size_t)(3 * warp) * KC;
                const double* xs = st + (size_t)ROWS * KC;
#pragma unroll
                for (int k = 2 * lane; k < KC; k += 64) {
                    const double2 a0 = *reinterpret_cast<const double2*>(q0 + k);
                    const double2 a1 = *reinterpret_cast<const double2*>(q0 + KC + k);
                    const double2 a2 = *reinterpret_cast<const double2*>(q0 + 2 * KC + k);
#pragma unroll
                    for (int j = 0; j < RP; ++j) {
                        const double2 x = *reinterpret_cast<const double2*>(xs + (size_t)j * KC + k);
                        acc[0][j] = fma(a0.x, x.x, acc[0][j]); acc[0][j] = fma(a0.y, x.y, acc[0][j]);
                        acc[1][j] = fma(a1.x, x.x, acc[1][j]); acc[1][j] = fma(a1.y, x.y, acc[1][j]);
                        acc[2][j] = fma(a2.x, x.x, acc[2][j]); acc[2][j] = fma(a2.y, x.y, acc[2][j]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (warp < nb) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < RP; ++j) { double v = wsum(acc[a][j]); if (lane == 0) p.out[(size_t)(3 * (b0 + warp) + a) * RP + j] = v; }
        }
    }
}

typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled_t get_encode() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    return (EncodeTiled_t)fn;
}
static CUtensorMap make_map(EncodeTiled_t enc, const double* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems, uint32_t box_cols, uint32_t box_rows,
                            CUtensorMapL2promotion l2) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {pitch_elems * 8}; cuuint32_t box[2] = {box_cols, box_rows}; cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

// V_pat: access-pattern probe — the k_cam addressing (warp per camera, 3 rows, 512 B per row per step) with no operand
// and no FMAs: separates the DRAM-pattern cost from the in-SM cost.
template <int UNROLL>
__global__ void __launch_bounds__(512, 1) k_pat(const P p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lo = (int)((long long)blockIdx.x * p.N / p.G), hi = (int)((long long)(blockIdx.x + 1) * p.N / p.G);
    const size_t ldq = p.ldq;
    double s = 0;
    for (int cam = lo + warp; cam < hi; cam += 16) {
        const double* q0 = p.Q + (size_t)(3 * cam) * ldq;
#pragma unroll UNROLL
        for (int k = 2 * lane; k < p.ldq; k += 64) {
            double2 a0 = ld_na(q0 + k), a1 = ld_na(q0 + ldq + k), a2 = ld_na(q0 + 2 * ldq + k);
            s += (a0.x + a0.y) + (a1.x + a1.y) + (a2.x + a2.y);
        }
    }
    if (s == 1.2345e300) p.out[0] = s;
}

// ---------------------------------------------------------------------------------------------------------------
template <typename F>
float time_it(F launch, int iters) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters;
}

template <int RP>
void run_all(int N, int iters) {
    const int n3 = 3 * N, ldq = (n3 + 63) / 64 * 64, r = RP;
    std::vector<double> Qh((size_t)n3 * ldq, 0.0), Xh((size_t)r * ldq, 0.0), ref((size_t)n3 * r, 0.0);
    srand(1);
    for (int i = 0; i < n3; ++i) for (int k = 0; k < n3; ++k) Qh[(size_t)i * ldq + k] = (rand() / (double)RAND_MAX) - 0.5;
    for (int j = 0; j < r; ++j) for (int k = 0; k < n3; ++k) Xh[(size_t)j * ldq + k] = (rand() / (double)RAND_MAX) - 0.5;
    const int check_rows = 64;
    for (int t = 0; t < check_rows; ++t) { int i = (int)((long long)t * (n3 - 1) / (check_rows - 1));
        for (int j = 0; j < r; ++j) { double s = 0; for (int k = 0; k < n3; ++k) s += Qh[(size_t)i * ldq + k] * Xh[(size_t)j * ldq + k]; ref[(size_t)i * r + j] = s; } }
    double *Q, *Xt, *out; int* err;
    CK(cudaMalloc(&Q, Qh.size() * 8)); CK(cudaMalloc(&Xt, Xh.size() * 8)); CK(cudaMalloc(&out, (size_t)n3 * r * 8)); CK(cudaMalloc(&err, 4));
    CK(cudaMemcpy(Q, Qh.data(), Qh.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(Xt, Xh.data(), Xh.size() * 8, cudaMemcpyHostToDevice));
    int dev = 0, sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const double bytes = 8.0 * n3 * (double)n3 + 16.0 * n3 * r;
    std::vector<double> oh((size_t)n3 * r);
    auto report = [&](const char* name, float ms) {
        CK(cudaMemcpy(oh.data(), out, oh.size() * 8, cudaMemcpyDeviceToHost));
        double e = 0;
        for (int t = 0; t < check_rows; ++t) { int i = (int)((long long)t * (n3 - 1) / (check_rows - 1)); for (int j = 0; j < r; ++j) e = fmax(e, fabs(oh[(size_t)i * r + j] - ref[(size_t)i * r + j])); }
        printf("%-44s %8.2f us  %8.1f GB/s   maxerr %.2e\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9, e);
        fflush(stdout);
        CK(cudaMemset(out, 0, oh.size() * 8));
    };
    printf("# N=%d n3=%d ldq=%d r=%d  Q=%.1f MB  SMs=%d\n", N, n3, ldq, r, 8.0 * n3 * ldq / 1e6, sms);
    P p{Q, Xt, out, N, n3, ldq, r, sms, 1};
    { float ms = time_it([&] { k_read<<<sms * 2, 512>>>(p); }, iters); printf("%-44s %8.2f us  %8.1f GB/s (pure read of Q)\n", "read-probe 2 CTA/SM", ms * 1e3, 8.0 * n3 * ldq / (ms * 1e-3) / 1e9); }
    { float ms = time_it([&] { k_read<<<sms * 4, 512>>>(p); }, iters); printf("%-44s %8.2f us  %8.1f GB/s (pure read of Q)\n", "read-probe 4 CTA/SM", ms * 1e3, 8.0 * n3 * ldq / (ms * 1e-3) / 1e9); }
#define CAM(UN, LD, NT, KSV, G_, label) { p.KS = KSV; p.G = G_; char nm[96]; snprintf(nm, 96, "cam %s un%d NT%d KS%d G%d", label, UN, NT, KSV, G_); \
        float ms = time_it([&] { k_cam<RP, UN, LD, NT><<<G_, NT>>>(p); }, iters); report(nm, ms); }
    CAM(4, 0, 512, 1, sms, "na");
    CAM(8, 0, 512, 1, sms, "na");
    CAM(2, 0, 512, 1, sms, "na");
    CAM(4, 1, 512, 1, sms, "plain");
    CAM(4, 2, 512, 1, sms, "cs");
    CAM(4, 0, 512, 2, sms, "na");
    CAM(4, 0, 512, 4, sms, "na");
    CAM(4, 0, 512, 8, sms, "na");
    CAM(4, 0, 512, 16, sms, "na");
    CAM(8, 0, 512, 4, sms, "na");
    CAM(4, 0, 1024, 1, sms, "na");
    CAM(4, 0, 1024, 2, sms, "na");
    CAM(4, 0, 1024, 4, sms, "na");
    CAM(4, 0, 256, 1, 2 * sms, "na");
    CAM(4, 0, 256, 2, 2 * sms, "na");
    CAM(4, 0, 512, 1, 2 * sms, "na");
    CAM(4, 0, 512, 2, 2 * sms, "na");
    { p.G = sms; float ms = time_it([&] { k_row<RP, 4><<<sms, 512>>>(p); }, iters); report("row-per-warp un4 G148", ms); }
    { float ms = time_it([&] { k_row<RP, 8><<<sms, 512>>>(p); }, iters); report("row-per-warp un8 G148", ms); }
    { float ms = time_it([&] { k_row<RP, 4><<<2 * sms, 512>>>(p); }, iters); report("row-per-warp un4 G296", ms); }
    { p.G = sms; float ms = time_it([&] { k_pat<4><<<sms, 512>>>(p); }, iters); printf("%-44s %8.2f us  %8.1f GB/s (pattern probe, no math)\n", "pat un4", ms * 1e3, 8.0 * n3 * ldq / (ms * 1e-3) / 1e9); }
    { p.G = sms; float ms = time_it([&] { k_pat<8><<<sms, 512>>>(p); }, iters); printf("%-44s %8.2f us  %8.1f GB/s (pattern probe, no math)\n", "pat un8", ms * 1e3, 8.0 * n3 * ldq / (ms * 1e-3) / 1e9); }
    { p.G = sms; float ms = time_it([&] { k_pat<16><<<sms, 512>>>(p); }, iters); printf("%-44s %8.2f us  %8.1f GB/s (pattern probe, no math)\n", "pat un16", ms * 1e3, 8.0 * n3 * ldq / (ms * 1e-3) / 1e9); }
    const int cpc = (N + sms - 1) / sms;
    auto tma2 = [&](auto kern, int NWC, int KC, int ST) {
        size_t sm = (size_t)ST * (3 * NWC + RP) * KC * 8 + 2 * ST * 8 + 128;
        if (sm > 227 * 1024) return;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); CK(cudaMemset(err, 0, 4));
        p.G = sms; p.KS = 1;
        float ms = time_it([&] { kern<<<sms, 32 * (NWC + 1), sm>>>(p, KC, ST, err); }, iters);
        int he; CK(cudaMemcpy(&he, err, 4, cudaMemcpyDeviceToHost));
        char nm[96]; snprintf(nm, 96, "tma2 NWC%d KC%d stages%d smem%zuKB err%d", NWC, KC, ST, sm / 1024, he); report(nm, ms);
    };
    if (cpc <= 12) {
        for (int KC : {256}) for (int ST : {2}) tma2(k_tma2<RP, 12>, 12, KC, ST);
    }
    EncodeTiled_t enc = get_encode();
    auto tma3 = [&](auto kern, int NWC, int KC, int ST, CUtensorMapL2promotion l2, const char* l2n) {
        size_t sm = (size_t)ST * (3 * NWC + RP) * KC * 8 + 2 * ST * 8 + 128;
        if (sm > 227 * 1024) return;
        CUtensorMap mq = make_map(enc, Q, ldq, n3, ldq, KC, 3, l2), mx = make_map(enc, Xt, ldq, RP, ldq, KC, RP, l2);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); CK(cudaMemset(err, 0, 4));
        p.G = sms; p.KS = 1;
        float ms = time_it([&] { kern<<<sms, 32 * (NWC + 1), sm>>>(p, mq, mx, ST, err); }, iters);
        int he; CK(cudaMemcpy(&he, err, 4, cudaMemcpyDeviceToHost));
        char nm[96]; snprintf(nm, 96, "tma3 NWC%d KC%d st%d %s smem%zuKB err%d", NWC, KC, ST, l2n, sm / 1024, he); report(nm, ms);
    };
#define T3(NWC, KC) for (int ST : {2, 3, 4, 5, 6, 8}) { tma3(k_tma3<RP, NWC, KC>, NWC, KC, ST, CU_TENSOR_MAP_L2_PROMOTION_NONE, "l2none"); \
                                                         tma3(k_tma3<RP, NWC, KC>, NWC, KC, ST, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "l2_256"); }
    if (cpc <= 12) { T3(12, 64) T3(12, 128) T3(12, 192) T3(12, 256) }
    T3(15, 128) T3(15, 192) T3(6, 128) T3(6, 256)
    cudaFree(Q); cudaFree(Xt); cudaFree(out); cudaFree(err);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1723, r = argc > 2 ? atoi(argv[2]) : 3, iters = argc > 3 ? atoi(argv[3]) : 50;
    if (r == 3) run_all<3>(N, iters); else if (r == 5) run_all<5>(N, iters); else if (r == 10) run_all<10>(N, iters); else printf("r must be 3, 5 or 10\n");
    return 0;
}
