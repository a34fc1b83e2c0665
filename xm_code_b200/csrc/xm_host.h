// xm_host.h — internal: the handle behind the C-ABI (shared by xm_capi.cu and xm_certify.cu).
#pragma once
#include "../../include/xm_b200.h"
#include "xm_device.cuh"
#include <string>

struct xm_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    xm_options opt{};
    std::string err;
    int num_sm = 0;
    // operator
    int N = 0, n3 = 0, ldq = 0;
    double* Qp = nullptr; size_t Qp_cap = 0;          // padded row-major dense Q
    double* Qstage = nullptr; size_t Qstage_cap = 0;   // staging for the user's column-major matrix
    int* bsr_rowptr = nullptr; int* bsr_col = nullptr; double* bsr_val = nullptr; int bsr_bdim = 0; bool is_bsr = false;
    // boundary-only operand exchange of a block-CSR operator on a communicator (xm_set_q_bsr): see Dev::peer_mask / need_cams
    unsigned char* d_peer_mask = nullptr; int* d_need_cams = nullptr; int n_need = 0; long long halo_sent = 0;
    // workspace (one allocation, carved per (N, r))
    char* ws = nullptr; size_t ws_cap = 0; int ws_r = -1, ws_N = -1, ws_G = -1, ws_ldq = -1, ws_bsr = -1;
    xm::Dev dev{};                                         // pointer template, filled by carve()
    // small persistent device objects
    xm::DevStats* d_stats = nullptr; xm::LogRec* d_log = nullptr; int* d_abort = nullptr;
    unsigned long long* d_bar = nullptr;               // 256 B: [0] barrier counter (single GPU), [16] epoch carried across launches
    double* d_scalar = nullptr;
    // I/O staging in the wire layout (device)
    double *io_R0 = nullptr, *io_s0 = nullptr, *io_v = nullptr, *io_Rout = nullptr, *io_sout = nullptr, *io_P = nullptr, *io_ps = nullptr;
    size_t io_cap_R = 0, io_cap_s = 0;
    // pinned host mirrors
    xm::DevStats* h_stats = nullptr; xm::LogRec* h_log = nullptr;
    int launches = 0;
    // cuBLAS / cuSOLVER handles of the certificate and the assembly (plain library GEMM / Cholesky / small eigen-solves), created on
    // first use and kept: creating and destroying them per call costs tens of milliseconds and device-wide synchronisations
    void* cublas = nullptr; void* cusolver = nullptr;
    // TMA descriptors (2-D tensor maps) for the dense Q and the current operand buffer
    CUtensorMap mapQ[3]{}, mapX{};
    void* encode_tiled = nullptr;       // cuTensorMapEncodeTiled, resolved through the runtime (no libcuda link)
    int smem_optin = 0;                 // max opt-in dynamic shared memory per block
    // operator rows held by this handle: cameras [cam0, cam1) (all of them unless a communicator is attached)
    int cam0 = 0, cam1 = 0;
    // multi-GPU communicator (xm_comm_*): one peer-mapped arena per rank with an identical layout everywhere
    //   [bar 256][abort 256][partials][ll: barrier inbox][slots: tagged partial sums][Xt: max_r * ldq][XtLL: tagged operand staging][outR][outS]
    int world = 1, rank = 0, comm_G = 0, comm_N = 0, comm_maxr = 0;
    bool comm_connected = false, comm_broken = false;
    char* arena = nullptr; size_t arena_bytes = 0;
    char* peer_arena[xm::kMaxWorld] = {};
    bool peer_ipc[xm::kMaxWorld] = {};      // opened with cudaIpcOpenMemHandle (to be closed)
    size_t off_bar = 0, off_abort = 0, off_partials = 0, off_ll = 0, off_slots = 0, off_xt = 0, off_xtll = 0, off_outR = 0, off_outS = 0;
};

// NVTX range around a C-ABI call (header-only NVTX v3: a no-op unless a profiler is attached)
#include <nvtx3/nvToolsExt.h>
struct XmRange {
    explicit XmRange(const char* name) { nvtxRangePushA(name); }
    ~XmRange() { nvtxRangePop(); }
};

#define XM_CUDA(h, call)                                                                           \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return XM_ECUDA;                                                                       \
        }                                                                                          \
    } while (0)

