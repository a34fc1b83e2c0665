// xm_inst.cu — instantiations of the persistent solve kernel and the op-level kernel for one group of padded ranks and ONE
// Q.Y path.  Compiled nine times (-DXM_INST_GROUP=0/1/2 x -DXM_INST_PATH=0/1/2) so the build runs in parallel; see
// xm_capi.cu:launch_any.  PATH: 0 = dense through the TMA ring, 1 = dense by direct loads, 2 = block-CSR.
#include "xm_host.h"
#include "xm_solve.cuh"

using namespace xm;

#if !defined(XM_INST_GROUP) || !defined(XM_INST_PATH) || !defined(XM_INST_MG)
#error "compile with -DXM_INST_GROUP=0|1|2 -DXM_INST_PATH=0|1|2 -DXM_INST_MG=0|1"
#endif
constexpr int kPath = XM_INST_PATH;
constexpr bool kMG = (XM_INST_MG != 0);        // 0: the one-GPU instantiation (multi-GPU branches folded away); 1: what a communicator launches
#define XM_CAT_(g, p, m) xm_launch_group##g##_path##p##_mg##m
#define XM_GROUP_FN2(g, p, m) XM_CAT_(g, p, m)
#define XM_GROUP_FN(g, p) XM_GROUP_FN2(g, p, XM_INST_MG)

template <int RP, int NT>
static cudaError_t launch_t(int kind, const xm_handle* h, const Dev& d, int opcode, size_t dyn, cudaStream_t st) {
    const void* fn = (kind == 0) ? (const void*)xm_solve_kernel<RP, NT, kPath, kMG> : (const void*)xm_ops_kernel<RP, NT, kPath, kMG>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    if (kind == 0) {
        void* args[] = {(void*)&d, (void*)h->mapQ, (void*)&h->mapX};
        return cudaLaunchCooperativeKernel(fn, dim3(d.G), dim3(NT), args, dyn, st);
    }
    void* args[] = {(void*)&d, (void*)h->mapQ, (void*)&h->mapX, (void*)&opcode};
    return cudaLaunchCooperativeKernel(fn, dim3(d.G), dim3(NT), args, dyn, st);
}

// block-CSR kernels hold 3 accumulators per lane whatever the rank: the same thread count for every RP — 1024 (the default: resident
// warps hide the operand gather's latency) or 512 (A/B hook XM_TUNE_BSR_NT); the plan decides (d.NW)
template <int RP, int NT_DENSE>
static cudaError_t launch_rp(int kind, const xm_handle* h, const Dev& d, int opcode, size_t dyn, cudaStream_t st) {
#if XM_INST_PATH == 2
    if (d.NW == 32) return launch_t<RP, 1024>(kind, h, d, opcode, dyn, st);
    return launch_t<RP, 512>(kind, h, d, opcode, dyn, st);
#else
    if (NT_DENSE == 256 && RP <= 10 && d.NW == 16) return launch_t<RP, (RP <= 10 ? 512 : 256)>(kind, h, d, opcode, dyn, st);   // A/B hook: one camera per warp
    return launch_t<RP, NT_DENSE>(kind, h, d, opcode, dyn, st);
#endif
}

#if XM_INST_GROUP == 0
cudaError_t XM_GROUP_FN(0, XM_INST_PATH)(int kind, int RP, const xm_handle* h, const Dev& d, int opcode, size_t dyn, cudaStream_t st) {
    switch (RP) {
        case 3: return launch_rp<3, 512>(kind, h, d, opcode, dyn, st);
        case 4: return launch_rp<4, 512>(kind, h, d, opcode, dyn, st);
        case 5: return launch_rp<5, 512>(kind, h, d, opcode, dyn, st);
    }
    return cudaErrorInvalidValue;
}
#elif XM_INST_GROUP == 1
cudaError_t XM_GROUP_FN(1, XM_INST_PATH)(int kind, int RP, const xm_handle* h, const Dev& d, int opcode, size_t dyn, cudaStream_t st) {
    switch (RP) {
        case 6: return launch_rp<6, 512>(kind, h, d, opcode, dyn, st);
        case 8: return launch_rp<8, 256>(kind, h, d, opcode, dyn, st);
        case 10: return launch_rp<10, 256>(kind, h, d, opcode, dyn, st);
    }
    return cudaErrorInvalidValue;
}
#else
cudaError_t XM_GROUP_FN(2, XM_INST_PATH)(int kind, int RP, const xm_handle* h, const Dev& d, int opcode, size_t dyn, cudaStream_t st) {
    switch (RP) {
        case 12: return launch_rp<12, 256>(kind, h, d, opcode, dyn, st);
        case 16: return launch_rp<16, 256>(kind, h, d, opcode, dyn, st);
        case 20: return launch_rp<20, 256>(kind, h, d, opcode, dyn, st);
    }
    return cudaErrorInvalidValue;
}
#endif
