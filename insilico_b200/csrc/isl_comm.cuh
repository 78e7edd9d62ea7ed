// =============================================================================
// isl_comm.cuh -- interface-row exchange between the engines of one box, inside the engine:
// one process per GPU, element blocks per GPU, owned row ranges; after the local assembly every rank sends the
// values of its ghost rows (rows it assembles into but another rank owns) to the owner, which adds them.
// Included by isl_engine.cu.  The reference has no counterpart (its only parallel construct is the OpenMP loop of
// base/auxi/parallel.hpp:25-60); this replaces the torch.distributed point-to-point calls of round 1, so a C++ caller
// (the binding) can drive several GPUs, and lets the send overlap the interior work.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already loaded by torch when there is one, else the
// system library), so the engine library itself has no link-time dependency and loads on machines without NCCL.
//
// Plan (built once per pattern by isl_exchange_setup, collective):
//   every rank makes the (global row, global column) key of every entry of its ghost rows and sends them to the
//   owner; the owner finds the position of every key in its own CSR (k_comm_locate) and keeps the position list.
// Step (isl_exchange): the ghost values go out as ONE contiguous slice of the value array per owner (the local
// numbering puts the ghost rows of one owner next to each other), the rhs rows likewise; the owner adds what it
// receives with k_unpack_add.  The patches of the Q1 row kernel that own INTERFACE rows are launched on a second,
// high-priority stream and the exchange is queued right behind them on that stream, while the interior patches run
// on the engine stream at the same time (the two sets write disjoint rows); comm_join() brings the streams together
// (launch_q1 in isl_engine.cu).  Other kernels (generic elements, general Q1 elements) exchange after their launch.
// =============================================================================
#pragma once

#include <dlfcn.h>

namespace nccl_dyn {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };   // nccl.h ncclDataType_t
typedef int (*fn_GetUniqueId)(ncclUniqueId*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_ErrStr)(int);
struct Api {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr; fn_CommInitRank CommInitRank = nullptr; fn_CommDestroy CommDestroy = nullptr;
    fn_Send Send = nullptr; fn_Recv Recv = nullptr; fn_AllGather AllGather = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr; fn_ErrStr GetErrorString = nullptr;
};
inline Api& api() {
    static Api a;
    if (a.lib) return a;
    void* l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // torch's copy when it is in the process
    if (!l) l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!l) throw IslError(std::string("NCCL library not found (libnccl.so.2): ") + dlerror());
    auto sym = [&](const char* n) { void* p = dlsym(l, n); if (!p) throw IslError(std::string("NCCL symbol missing: ") + n); return p; };
    a.GetUniqueId = (fn_GetUniqueId)sym("ncclGetUniqueId"); a.CommInitRank = (fn_CommInitRank)sym("ncclCommInitRank");
    a.CommDestroy = (fn_CommDestroy)sym("ncclCommDestroy"); a.Send = (fn_Send)sym("ncclSend"); a.Recv = (fn_Recv)sym("ncclRecv");
    a.AllGather = (fn_AllGather)sym("ncclAllGather"); a.GroupStart = (fn_Group)sym("ncclGroupStart");
    a.GroupEnd = (fn_Group)sym("ncclGroupEnd"); a.GetErrorString = (fn_ErrStr)sym("ncclGetErrorString");
    a.lib = l;
    return a;
}
}  // namespace nccl_dyn

#define ISL_NCCL(call)                                                                                                  \
    do {                                                                                                               \
        int r__ = (call);                                                                                              \
        if (r__ != 0) throw IslError(std::string("NCCL error: ") + nccl_dyn::api().GetErrorString(r__) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

struct CommSend { int dst; int64_t val_lo, n_val, row_lo, n_rows; };
struct CommRecv { int src; int64_t n_val, n_rows; DevBuf<int64_t> pos, rows; DevBuf<double> bval, brhs; };

struct CommState {
    nccl_dyn::ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_iface = nullptr;
    bool plan = false, iface_event_valid = false, join_pending = false;
    std::vector<CommSend> sends;
    std::vector<std::unique_ptr<CommRecv>> recvs;
    DevBuf<uint8_t> iface_row;   // [n_local] 1: the row is sent to or received from another rank
    int64_t plan_nnz = -1;
    ~CommState() {
        if (comm) nccl_dyn::api().CommDestroy(comm);
        if (ev_ready) cudaEventDestroy(ev_ready);
        if (ev_done) cudaEventDestroy(ev_done);
        if (ev_iface) cudaEventDestroy(ev_iface);
        if (stream) cudaStreamDestroy(stream);
    }
};

// (global row, global column) keys of the entries of the local rows [lo, hi)
__global__ void k_comm_keys(const int64_t* rowptr, const int32_t* col, const int64_t* l2g, int64_t lo, int64_t hi, uint64_t* keys) {
    const int64_t base = rowptr[lo];
    for (int64_t r = lo + blockIdx.x; r < hi; r += gridDim.x) {
        const uint64_t gr = (uint64_t)l2g[r];
        for (int64_t s = rowptr[r] + threadIdx.x; s < rowptr[r + 1]; s += blockDim.x)
            keys[s - base] = (gr << 32) | (uint64_t)(uint32_t)l2g[col[s]];
    }
}
// global equation -> local row inside the owned range [own_lo, own_hi) (l2g ascending there), -1 when absent
__device__ __forceinline__ int64_t comm_find_owned(const int64_t* l2g, int64_t own_lo, int64_t own_hi, int64_t g) {
    int64_t lo = own_lo, hi = own_hi;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (l2g[mid] < g) lo = mid + 1; else hi = mid; }
    return (lo < own_hi && l2g[lo] == g) ? lo : -1;
}
// owner side: position of every received key in the local CSR (columns are compared through l2g: the local column
// order is not the global one)
__global__ void k_comm_locate(const uint64_t* keys, int64_t n, const int64_t* l2g, int64_t own_lo, int64_t own_hi,
                              const int64_t* rowptr, const int32_t* col, int64_t* pos, int* err) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gr = (int64_t)(keys[k] >> 32), gc = (int64_t)(keys[k] & 0xffffffffu);
        const int64_t r = comm_find_owned(l2g, own_lo, own_hi, gr);
        int64_t p = -1;
        if (r >= 0)
            for (int64_t s = rowptr[r]; s < rowptr[r + 1]; s++)
                if (l2g[col[s]] == gc) { p = s; break; }
        if (p < 0) { *err = 1; p = 0; }
        pos[k] = p;
    }
}
__global__ void k_comm_locate_rows(const int64_t* grows, int64_t n, const int64_t* l2g, int64_t own_lo, int64_t own_hi, int64_t* rows,
                                   uint8_t* iface, int* err) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = comm_find_owned(l2g, own_lo, own_hi, grows[k]);
        if (r < 0) { *err = 1; r = own_lo; }
        rows[k] = r;
        iface[r] = 1;
    }
}
__global__ void k_comm_mark(uint8_t* iface, int64_t lo, int64_t hi) {
    for (int64_t r = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += (int64_t)gridDim.x * blockDim.x) iface[r] = 1;
}
// patches that own an interface row come first in the launch order
__global__ void k_comm_patch_flags(const int32_t* p_row_off, const int32_t* rows, const uint8_t* iface, int n_patches, int32_t* flag) {
    const int pid = blockIdx.x;
    if (pid >= n_patches) return;
    int f = 0;
    for (int r = p_row_off[pid] + threadIdx.x; r < p_row_off[pid + 1]; r += blockDim.x) f |= iface[rows[r]];
    f = __syncthreads_or(f);
    if (threadIdx.x == 0) flag[pid] = f ? 1 : 0;
}
