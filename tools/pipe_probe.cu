// tools/pipe_probe.cu — measures raw SM pipe rates on B200 that bound the Q.Y consumer loop:
//   DFMA throughput (lane-FMAs / clk / SM), LDS.128 throughput (B / clk / SM), and both together.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k_dfma(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512, 1) k_lds(double* out, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 s = make_double2(0, 0);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double2 v = *reinterpret_cast<const double2*>(sm + ((warp * 512 + u * 64 + 2 * lane + i * 64) & 16383));
            s.x += v.x; s.y += v.y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s.x + s.y;
}

// the consumer step of the Q.Y ring: 6 LDS.128 + 18 DFMA per 64 columns (r = 3)
__global__ void __launch_bounds__(512, 1) k_step(double* out, int iters, int nwarps_active) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 20480; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= nwarps_active) return;
    double acc[3][3] = {};
    const double* q0 = sm + warp * 3 * 128;
    const double* xs = sm + 16 * 3 * 128;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k0 = 0; k0 < 128; k0 += 64) {
            const int k = k0 + 2 * lane;
            const double2 a0 = *reinterpret_cast<const double2*>(q0 + k);
            const double2 a1 = *reinterpret_cast<const double2*>(q0 + 128 + k);
            const double2 a2 = *reinterpret_cast<const double2*>(q0 + 256 + k);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double2 x = *reinterpret_cast<const double2*>(xs + j * 128 + k);
                acc[0][j] = fma(a0.x, x.x, acc[0][j]); acc[0][j] = fma(a0.y, x.y, acc[0][j]);
                acc[1][j] = fma(a1.x, x.x, acc[1][j]); acc[1][j] = fma(a1.y, x.y, acc[1][j]);
                acc[2][j] = fma(a2.x, x.x, acc[2][j]); acc[2][j] = fma(a2.y, x.y, acc[2][j]);
            }
        }
    }
    double s = 0;
    for (int a = 0; a < 3; ++a) for (int j = 0; j < 3; ++j) s += acc[a][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    int sms = 0, khz = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0)); CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    const double ghz = khz / 1e6;
    printf("SMs %d, clock %.3f GHz (attr)\n", sms, ghz);
    for (int nt : {128, 256, 512}) {
        const int iters = 20000;
        float ms = timeit([&] { k_dfma<8><<<sms, nt>>>(out, iters, 1.0000001, 1e-9); });
        double fmas = (double)sms * nt * 8.0 * iters;
        printf("DFMA  threads/SM %4d : %.2f TFLOP/s  = %.1f lane-FMA/clk/SM\n", nt, 2 * fmas / (ms * 1e-3) / 1e12, fmas / (ms * 1e-3) / sms / (ghz * 1e9));
    }
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    for (int nt : {128, 256, 512}) {
        const int iters = 4000;
        float ms = timeit([&] { k_lds<<<sms, nt, 16384 * 8>>>(out, iters); });
        double bytes = (double)sms * nt * 8.0 * iters * 16;
        printf("LDS.128 threads/SM %4d : %.1f B/clk/SM\n", nt, bytes / (ms * 1e-3) / sms / (ghz * 1e9));
    }
    CK(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480 * 8));
    for (int nw : {1, 4, 8, 12, 15, 16}) {
        const int iters = 20000;
        float ms = timeit([&] { k_step<<<sms, 512, 20480 * 8>>>(out, iters, nw); });
        printf("Q.Y consumer step, %2d warps/SM: %.3f us per [nw x 3 rows x 128 cols] chunk  (%.0f cycles)\n", nw, ms * 1e3 / iters, ms * 1e-3 / iters * ghz * 1e9);
    }
    return 0;
}
