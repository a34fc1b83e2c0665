// tools/bsr_tune.cu — standalone micro-benchmark of block-CSR FP64 Q.Y variants at Erdos-Renyi scale (tuning tool, not shipped).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/bsr_tune.cu -o tools/bsr_tune
// Run  : tools/bsr_tune [N_cameras=100000] [degree=100] [r=10] [iters=20]
//
// Question it answers (profiles/r01_bsr_qy.md): the shipped kernel (one 512-thread CTA per SM inside the persistent solver,
// per-warp bulk-TMA staging of 32-block chunks, 4 operand gathers per sub-warp in flight) reaches 32 % of the HBM line at
// r = 10 with no unit saturated — is that occupancy?  Variants differ ONLY in resident warps and gather depth:
//   stage<NT, K>   per-warp double-buffered staging (one cp.async.bulk per 32-block chunk), NT threads per CTA, one CTA per SM
//   direct<NT, K>  blocks through registers (three 256-bit loads per block), no staging, 1024 / NT CTAs per SM (32 warps)
// Same data layout as the library: blocks 4x4 row-major padded (128 B), int32 columns, operand camera-major X[(3c+a) r + j],
// result camera-major.  The graph is random (uniform columns, sorted per row) — the adversarial no-locality case.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct P {
    int N, r, W, cpw;                 // cameras, rank, sub-warp width (power of two >= r), sub-warps per warp
    int xs;                           // operand doubles per camera: 3r, or 3r rounded up to a multiple of 4 (32-byte sectors never straddled)
    const int* rowptr; const int* col; const double* val; const double* X; double* out;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldg_v4(const double* p, double (&v)[4]) {
    asm("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ double2 lds_v2(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double shfl_xor_d(double v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }

// reference: one thread per (row, a, j)
__global__ void k_ref(const P p) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.N * 3 * p.r) return;
    const int j = (int)(t % p.r), a = (int)((t / p.r) % 3), i = (int)(t / (3 * p.r));
    double acc = 0.0;
    for (int b = p.rowptr[i]; b < p.rowptr[i + 1]; ++b) {
        const int c = p.col[b];
        const double* q = p.val + (size_t)b * 16 + 4 * a;
        acc += q[0] * p.X[(size_t)c * p.xs + j] + q[1] * p.X[(size_t)c * p.xs + p.r + j] + q[2] * p.X[(size_t)c * p.xs + 2 * p.r + j];
    }
    p.out[(size_t)(3 * i + a) * p.r + j] = acc;
}

// ---- direct: warp per row, rows strided over all warps of the grid; sub-warp (block, column) mapping; K gathers in flight
template <int NT, int K>
__global__ void __launch_bounds__(NT, 1024 / NT) k_direct(const P p) {      // 32 resident warps per SM, <= 64 registers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = p.W, cpw = p.cpw, sw = lane / W, j = lane % W, r = p.r;
    const bool act = j < r;
    const int nwarps = gridDim.x * (NT / 32);
    for (int row = blockIdx.x * (NT / 32) + warp; row < p.N; row += nwarps) {
        const int b0 = p.rowptr[row], b1 = p.rowptr[row + 1];
        double e0 = 0, e1 = 0, e2 = 0;
        for (int g0 = b0; g0 < b1; g0 += K * cpw) {
            double x[K][3], q0[K][4], q1[K][4], q2[K][4];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int b = g0 + k * cpw + sw;
                x[k][0] = x[k][1] = x[k][2] = 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t) q0[k][t] = q1[k][t] = q2[k][t] = 0.0;
                if (b < b1) {
                    const int c = __ldg(p.col + b);
                    const double* blk = p.val + (size_t)b * 16;
                    ldg_v4(blk, q0[k]); ldg_v4(blk + 4, q1[k]); ldg_v4(blk + 8, q2[k]);
                    if (act) { const double* xp = p.X + (size_t)c * p.xs + j; x[k][0] = xp[0]; x[k][1] = xp[r]; x[k][2] = xp[2 * r]; }
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                e0 = fma(q0[k][0], x[k][0], e0); e0 = fma(q0[k][1], x[k][1], e0); e0 = fma(q0[k][2], x[k][2], e0);
                e1 = fma(q1[k][0], x[k][0], e1); e1 = fma(q1[k][1], x[k][1], e1); e1 = fma(q1[k][2], x[k][2], e1);
                e2 = fma(q2[k][0], x[k][0], e2); e2 = fma(q2[k][1], x[k][1], e2); e2 = fma(q2[k][2], x[k][2], e2);
            }
        }
        for (int off = W; off < 32; off <<= 1) { e0 += shfl_xor_d(e0, off); e1 += shfl_xor_d(e1, off); e2 += shfl_xor_d(e2, off); }
        if (lane < W && act) {
            double* o = p.out + (size_t)(3 * row) * r + j;
            o[0] = e0; o[r] = e1; o[2 * r] = e2;
        }
    }
}

// ---- stage: like the library — per-warp double-buffered 32-block chunks through shared memory (cp.async.bulk + mbarrier)
template <int NT, int K, int CH>      // CH = blocks per staged chunk (<= 32: one column index per lane)
__global__ void __launch_bounds__(NT, 1) k_stage(const P p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = p.W, cpw = p.cpw, sw = lane / W, j = lane % W, r = p.r;
    const bool act = j < r;
    double* buf = reinterpret_cast<double*>(smem) + (size_t)warp * 2 * CH * 16;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(smem) + (size_t)NW * 2 * CH * 16) + warp * 2;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const int nwarps = gridDim.x * NW;
    // the warp's rows: row0, row0 + nwarps, ... ; chunks of CH blocks, one in flight ahead
    int row = blockIdx.x * NW + warp;
    int b0 = 0, b1 = 0, q = 0;
    auto seek = [&]() {                                   // first non-empty chunk at or after (row, q)
        while (row < p.N) {
            b0 = p.rowptr[row]; b1 = p.rowptr[row + 1];
            if (b0 + q * CH < b1) return true;
            row += nwarps; q = 0;
        }
        return false;
    };
    auto issue = [&](int rb0, int rb1, int qq, int bf) {
        const int st = rb0 + qq * CH, nb = min(CH, rb1 - st);
        __syncwarp();
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[bf])), "r"((unsigned)nb * 128u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(buf + (size_t)bf * CH * 16)), "l"(p.val + (size_t)st * 16), "r"((unsigned)nb * 128u), "r"(smem_u32(&bar[bf])), "l"(policy) : "memory");
        }
        return (lane < nb) ? __ldg(p.col + st + lane) : 0;
    };
    unsigned phase = 0;
    int bf = 0, colreg = 0;
    bool have = seek();
    if (have) colreg = issue(b0, b1, q, 0);
    double e0 = 0, e1 = 0, e2 = 0;
    while (have) {
        const int crow = row, cb0 = b0, cb1 = b1, cq = q;
        // next chunk
        q += 1;
        bool more = seek();
        int colnext = 0;
        if (more) colnext = issue(b0, b1, q, bf ^ 1);
        // consume the current one
        {
            unsigned ok = 0;
            while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar[bf])), "r"((phase >> bf) & 1u) : "memory");
            phase ^= 1u << bf;
        }
        const int st = cb0 + cq * CH, nb = min(CH, cb1 - st);
        const unsigned sb = smem_u32(buf + (size_t)bf * CH * 16);
        for (int g0 = 0; g0 < nb; g0 += K * cpw) {
            double x[K][3];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int bi = g0 + k * cpw + sw;
                const int c = __shfl_sync(0xffffffffu, colreg, bi & 31);
                x[k][0] = x[k][1] = x[k][2] = 0.0;
                if (act && bi < nb) { const double* xp = p.X + (size_t)c * p.xs + j; x[k][0] = xp[0]; x[k][1] = xp[r]; x[k][2] = xp[2 * r]; }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int bi = g0 + k * cpw + sw;
                if (bi < nb) {
                    const unsigned qa = sb + (unsigned)bi * 128u;
                    const double2 a0 = lds_v2(qa), a1 = lds_v2(qa + 16), c0 = lds_v2(qa + 32), c1 = lds_v2(qa + 48), d0 = lds_v2(qa + 64), d1 = lds_v2(qa + 80);
                    e0 = fma(a0.x, x[k][0], e0); e0 = fma(a0.y, x[k][1], e0); e0 = fma(a1.x, x[k][2], e0);
                    e1 = fma(c0.x, x[k][0], e1); e1 = fma(c0.y, x[k][1], e1); e1 = fma(c1.x, x[k][2], e1);
                    e2 = fma(d0.x, x[k][0], e2); e2 = fma(d0.y, x[k][1], e2); e2 = fma(d1.x, x[k][2], e2);
                }
            }
        }
        const bool row_done = !more || row != crow;
        if (row_done) {
            for (int off = W; off < 32; off <<= 1) { e0 += shfl_xor_d(e0, off); e1 += shfl_xor_d(e1, off); e2 += shfl_xor_d(e2, off); }
            if (lane < W && act) { double* o = p.out + (size_t)(3 * crow) * r + j; o[0] = e0; o[r] = e1; o[2 * r] = e2; }
            e0 = e1 = e2 = 0;
        }
        have = more; colreg = colnext; bf ^= 1;
    }
}

// ---- stage4: k_stage with the operand laid out [camera][column][4] (3 rows + pad = 32 bytes): ONE 256-bit gather per lane and block
template <int NT, int K, int CH>
__global__ void __launch_bounds__(NT, 1) k_stage4(const P p, const double* __restrict__ X4) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = p.W, cpw = p.cpw, sw = lane / W, j = lane % W, r = p.r;
    const bool act = j < r;
    double* buf = reinterpret_cast<double*>(smem) + (size_t)warp * 2 * CH * 16;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(smem) + (size_t)NW * 2 * CH * 16) + warp * 2;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const int nwarps = gridDim.x * NW;
    int row = blockIdx.x * NW + warp;
    int b0 = 0, b1 = 0, q = 0;
    auto seek = [&]() {
        while (row < p.N) {
            b0 = p.rowptr[row]; b1 = p.rowptr[row + 1];
            if (b0 + q * CH < b1) return true;
            row += nwarps; q = 0;
        }
        return false;
    };
    auto issue = [&](int rb0, int rb1, int qq, int bf) {
        const int st = rb0 + qq * CH, nb = min(CH, rb1 - st);
        __syncwarp();
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[bf])), "r"((unsigned)nb * 128u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(buf + (size_t)bf * CH * 16)), "l"(p.val + (size_t)st * 16), "r"((unsigned)nb * 128u), "r"(smem_u32(&bar[bf])), "l"(policy) : "memory");
        }
        return (lane < nb) ? __ldg(p.col + st + lane) : 0;
    };
    unsigned phase = 0;
    int bf = 0, colreg = 0;
    bool have = seek();
    if (have) colreg = issue(b0, b1, q, 0);
    double e0 = 0, e1 = 0, e2 = 0;
    while (have) {
        const int crow = row, cb0 = b0, cb1 = b1, cq = q;
        q += 1;
        bool more = seek();
        int colnext = 0;
        if (more) colnext = issue(b0, b1, q, bf ^ 1);
        {
            unsigned ok = 0;
            while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar[bf])), "r"((phase >> bf) & 1u) : "memory");
            phase ^= 1u << bf;
        }
        const int st = cb0 + cq * CH, nb = min(CH, cb1 - st);
        const unsigned sb = smem_u32(buf + (size_t)bf * CH * 16);
        for (int g0 = 0; g0 < nb; g0 += K * cpw) {
            double x[K][4];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int bi = g0 + k * cpw + sw;
                const int c = __shfl_sync(0xffffffffu, colreg, bi & 31);
                x[k][0] = x[k][1] = x[k][2] = x[k][3] = 0.0;
                if (act && bi < nb) {
                    const double* xp = X4 + ((size_t)c * r + j) * 4;
                    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x[k][0]), "=d"(x[k][1]), "=d"(x[k][2]), "=d"(x[k][3]) : "l"(xp));
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int bi = g0 + k * cpw + sw;
                if (bi < nb) {
                    const unsigned qa = sb + (unsigned)bi * 128u;
                    const double2 a0 = lds_v2(qa), a1 = lds_v2(qa + 16), c0 = lds_v2(qa + 32), c1 = lds_v2(qa + 48), d0 = lds_v2(qa + 64), d1 = lds_v2(qa + 80);
                    e0 = fma(a0.x, x[k][0], e0); e0 = fma(a0.y, x[k][1], e0); e0 = fma(a1.x, x[k][2], e0);
                    e1 = fma(c0.x, x[k][0], e1); e1 = fma(c0.y, x[k][1], e1); e1 = fma(c1.x, x[k][2], e1);
                    e2 = fma(d0.x, x[k][0], e2); e2 = fma(d0.y, x[k][1], e2); e2 = fma(d1.x, x[k][2], e2);
                }
            }
        }
        const bool row_done = !more || row != crow;
        if (row_done) {
            for (int off = W; off < 32; off <<= 1) { e0 += shfl_xor_d(e0, off); e1 += shfl_xor_d(e1, off); e2 += shfl_xor_d(e2, off); }
            if (lane < W && act) { double* o = p.out + (size_t)(3 * crow) * r + j; o[0] = e0; o[r] = e1; o[2 * r] = e2; }
            e0 = e1 = e2 = 0;
        }
        have = more; colreg = colnext; bf ^= 1;
    }
}

template <typename F>
static float time_ms(F launch, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / iters;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 100000, deg = argc > 2 ? atoi(argv[2]) : 100, r = argc > 3 ? atoi(argv[3]) : 10;
    const int iters = argc > 4 ? atoi(argv[4]) : 20;
    const int pad = argc > 5 ? atoi(argv[5]) : 0;          // 1: pad the operand's camera stride to a multiple of 32 bytes
    int W = 4; while (W < r) W <<= 1;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int SM = prop.multiProcessorCount;
    std::mt19937_64 rng(0);
    std::vector<int> rowptr(N + 1, 0), col;
    col.reserve((size_t)N * (deg + 1));
    for (int i = 0; i < N; ++i) {
        const int d = deg - 5 + (int)(rng() % 11);
        std::vector<int> c(d);
        for (int& v : c) v = (int)(rng() % N);
        c.push_back(i);
        std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end());
        col.insert(col.end(), c.begin(), c.end());
        rowptr[i + 1] = (int)col.size();
    }
    const size_t nnzb = col.size();
    const int xs = pad ? (3 * r + 3) / 4 * 4 : 3 * r;
    std::vector<double> val(nnzb * 16, 0.0), X((size_t)xs * N);
    std::uniform_real_distribution<double> U(-1, 1);
    for (size_t b = 0; b < nnzb; ++b) for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) val[b * 16 + a * 4 + c] = U(rng);
    for (double& v : X) v = U(rng);
    int *d_rowptr, *d_col; double *d_val, *d_X, *d_out, *d_ref;
    CK(cudaMalloc(&d_rowptr, sizeof(int) * (N + 1))); CK(cudaMalloc(&d_col, sizeof(int) * nnzb)); CK(cudaMalloc(&d_val, sizeof(double) * nnzb * 16));
    const size_t nout = (size_t)3 * N * r;
    CK(cudaMalloc(&d_X, sizeof(double) * X.size())); CK(cudaMalloc(&d_out, sizeof(double) * nout)); CK(cudaMalloc(&d_ref, sizeof(double) * nout));
    CK(cudaMemcpy(d_rowptr, rowptr.data(), sizeof(int) * (N + 1), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_col, col.data(), sizeof(int) * nnzb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_val, val.data(), sizeof(double) * nnzb * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_X, X.data(), sizeof(double) * X.size(), cudaMemcpyHostToDevice));
    P p{N, r, W, 32 / W, xs, d_rowptr, d_col, d_val, d_X, d_ref};
    const long long tot = (long long)N * 3 * r;
    k_ref<<<(unsigned)((tot + 255) / 256), 256>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<double> ref(nout), got(nout);
    CK(cudaMemcpy(ref.data(), d_ref, sizeof(double) * ref.size(), cudaMemcpyDeviceToHost));
    p.out = d_out;
    const double bytes = (double)nnzb * 132 + 4.0 * (N + 1) + 2.0 * 8 * 3 * N * r;
    printf("# N=%d nnzb=%zu r=%d W=%d operand stride %d doubles per camera  algorithmic bytes %.3f GB  SMs=%d\n", N, nnzb, r, W, xs, bytes / 1e9, SM);
    auto report = [&](const char* name, float ms) {
        CK(cudaMemcpy(got.data(), d_out, sizeof(double) * got.size(), cudaMemcpyDeviceToHost));
        double err = 0, mx = 0;
        for (size_t i = 0; i < got.size(); ++i) { err = std::max(err, std::abs(got[i] - ref[i])); mx = std::max(mx, std::abs(ref[i])); }
        printf("%-34s %9.3f us  %8.1f GB/s   rel.err %.2e\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9, err / mx);
        CK(cudaMemset(d_out, 0, sizeof(double) * got.size()));
    };
#define RUN_STAGE(NT, K, CH) do { \
        const size_t sm = (size_t)(NT / 32) * 2 * (CH * 128 + 8); \
        CK(cudaFuncSetAttribute(k_stage<NT, K, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        report("stage  NT=" #NT " K=" #K " CH=" #CH " 1 CTA/SM", time_ms([&]() { k_stage<NT, K, CH><<<SM, NT, sm>>>(p); }, iters)); } while (0)
#define RUN_DIRECT(NT, K) report("direct NT=" #NT " K=" #K " 1024/NT CTA/SM", time_ms([&]() { k_direct<NT, K><<<SM * (1024 / NT), NT>>>(p); }, iters))
    // operand in the [camera][column][4] layout for the stage4 variants
    std::vector<double> X4h((size_t)N * r * 4, 0.0);
    for (int c = 0; c < N; ++c) for (int jj = 0; jj < r; ++jj) for (int a = 0; a < 3; ++a) X4h[((size_t)c * r + jj) * 4 + a] = X[(size_t)c * xs + (size_t)a * r + jj];
    double* d_X4; CK(cudaMalloc(&d_X4, sizeof(double) * X4h.size())); CK(cudaMemcpy(d_X4, X4h.data(), sizeof(double) * X4h.size(), cudaMemcpyHostToDevice));
#define RUN_STAGE4(NT, K, CH) do { \
        const size_t sm = (size_t)(NT / 32) * 2 * (CH * 128 + 8); \
        CK(cudaFuncSetAttribute(k_stage4<NT, K, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        report("stage4 NT=" #NT " K=" #K " CH=" #CH " v4 operand", time_ms([&]() { k_stage4<NT, K, CH><<<SM, NT, sm>>>(p, d_X4); }, iters)); } while (0)
    RUN_STAGE4(1024, 2, 16);
    RUN_STAGE4(1024, 4, 16);
    RUN_STAGE4(512, 4, 32);
    RUN_STAGE(512, 4, 32);      // the library's configuration
    RUN_STAGE(512, 8, 32);
    RUN_STAGE(512, 4, 16);      // half-size chunks (isolates the chunk-size effect of the next two)
    RUN_STAGE(1024, 4, 16);     // twice the resident warps (<= 64 registers per thread); 16-block chunks so that 32 warps x 2 buffers fit
    RUN_STAGE(1024, 2, 16);
    RUN_DIRECT(512, 1);     // 2 CTAs x 512: 32 warps per SM, blocks through registers
    RUN_DIRECT(512, 2);
    RUN_DIRECT(1024, 1);
    RUN_DIRECT(256, 1);
    RUN_DIRECT(256, 2);
    return 0;
}
