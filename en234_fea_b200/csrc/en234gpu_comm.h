// Multi-GPU plumbing for element-partitioned runs: one process per GPU, NCCL over NVLink/NVSwitch.
// The reference is strictly serial (no communication backend); everything here is new.
//  * halo sum: interface (shared) nodes hold partial sums after a local operator application; each rank sends its
//    partials to every neighbour and adds what it receives (ncclSend/ncclRecv grouped; 1-D slabs => <= 2 neighbours)
//  * scalar all-reduces for the PCG dot products and norms (device-resident scalars, graph-capturable)
// NCCL is loaded with dlopen at communicator creation so that the library still loads on a host without it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "en234gpu_asm.cuh"
#include "en234gpu_p2p.cuh"

namespace en234k {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline std::string& comm_err_ref() { static thread_local std::string e; return e; }
inline const char* comm_error() { return comm_err_ref().c_str(); }

inline NcclApi* nccl_api() {
    static NcclApi api;
    if (api.lib) return &api;
    // torch bundles its own libnccl.so.2; if it is already loaded RTLD_NOLOAD finds it, otherwise use the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { comm_err_ref() = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
#define LD(field, name)                                                                   \
    api.field = (decltype(api.field))dlsym(h, name);                                      \
    if (!api.field) { comm_err_ref() = std::string("missing NCCL symbol ") + name; return nullptr; }
    LD(GetUniqueId, "ncclGetUniqueId") LD(CommInitRank, "ncclCommInitRank") LD(CommDestroy, "ncclCommDestroy")
    LD(AllReduce, "ncclAllReduce") LD(Send, "ncclSend") LD(Recv, "ncclRecv") LD(GroupStart, "ncclGroupStart")
    LD(GroupEnd, "ncclGroupEnd") LD(GetErrorString, "ncclGetErrorString")
#undef LD
    api.lib = h;
    return &api;
}

struct Comm {
    bool active = false;             // communicator created (n_ranks > 1)
    int n_ranks = 1, rank = 0;
    ncclComm_t nccl = nullptr;
    int n_nbr = 0;
    std::vector<int> nbr_rank, shared_ptr;    // host copies
    int* d_shared = nullptr;                  // device: local node index per shared entry (all neighbours concatenated)
    double *sendbuf = nullptr, *recvbuf = nullptr;
    int n_shared = 0;
    int launches_per_halo = 0;                // kernels of one halo sum (pack + one unpack per neighbour)
    int n_interface = 0;                      // interface nodes are local rows [0, n_interface) (0: not separable)
    // peer-memory path (en234gpu_p2p.cuh): arena shared through CUDA IPC, side stream for the overlapped halo exchange
    bool p2p = false;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    void* peer_base[P2P_MAXR] = {nullptr};
    PeerSet peers{};
    HaloSet halo{};
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_if = nullptr;
    // collectives folded into the PCG kernels (FusedComm, en234gpu_p2p.cuh): per interface node the (neighbour, position)
    // entries in ascending neighbour rank; EN234_P2P_FUSED=0 keeps the separate push / wait-add / all-reduce kernels
    bool fused = false;                       // interface tables built
    bool fused_loop = false;                  // PCG loop with the collectives folded into the kernels (EN234_P2P_FUSED=1)
    bool inproc = false;                      // ranks are threads of one process (en234gpu_create_multi): no NCCL at all
    double* d_scratch = nullptr;              // 4 doubles: operand of the peer-memory scalar all-reduces
    std::vector<int> h_shared;                // host copy of the shared-node lists
    int *d_if_ptr = nullptr, *d_if_nbr = nullptr, *d_if_pos = nullptr;
    unsigned char* d_if_before = nullptr;
    FusedComm F{};
};

constexpr size_t P2P_HDR_BYTES = (sizeof(ArenaHdr) + 255) / 256 * 256;

inline int comm_arena_create(Comm& c, int max_shared_nodes, char* handle64) {
    if (!c.active) { comm_err_ref() = "comm_arena_create: communicator not initialised"; return 1; }
    if (c.n_ranks > P2P_MAXR) { comm_err_ref() = "peer-memory path supports at most 16 ranks"; return 1; }
    c.peers.cap = 3LL * (max_shared_nodes > 0 ? max_shared_nodes : 1);
    c.arena_bytes = P2P_HDR_BYTES + (size_t)c.n_ranks * 2 * (size_t)c.peers.cap * sizeof(double);
    if (cudaMalloc(&c.arena, c.arena_bytes) != cudaSuccess) { comm_err_ref() = "cudaMalloc failed for the peer arena"; return 1; }
    cudaMemset(c.arena, 0, c.arena_bytes);
    if (!c.d_scratch && cudaMalloc((void**)&c.d_scratch, 4 * sizeof(double)) != cudaSuccess) { comm_err_ref() = "cudaMalloc failed"; return 1; }
    if (c.inproc) return 0;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c.arena);
    if (e != cudaSuccess) { comm_err_ref() = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return 1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    memcpy(handle64, &h, 64);
    return 0;
}

inline int comm_arena_finish(Comm& c);
// ranks in different processes: the peers' arenas are opened through their CUDA-IPC handles
inline int comm_arena_connect(Comm& c, const char* handles /* n_ranks x 64 */) {
    if (!c.arena) { comm_err_ref() = "comm_arena_connect: no arena"; return 1; }
    c.peers.n_ranks = c.n_ranks;
    c.peers.me = c.rank;
    for (int r = 0; r < c.n_ranks; ++r) {
        void* base = c.arena;
        if (r != c.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + 64 * (size_t)r, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { comm_err_ref() = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return 1; }
            c.peer_base[r] = base;
        }
        c.peers.arena[r] = (ArenaHdr*)base;
        c.peers.halo[r] = (double*)((char*)base + P2P_HDR_BYTES);
    }
    return comm_arena_finish(c);
}
// ranks are threads of one process: the peers' arenas are plain device pointers (peer access enabled by the caller)
inline int comm_arena_connect_inproc(Comm& c, void* const* bases /* n_ranks */) {
    if (!c.arena) { comm_err_ref() = "comm_arena_connect: no arena"; return 1; }
    c.peers.n_ranks = c.n_ranks;
    c.peers.me = c.rank;
    for (int r = 0; r < c.n_ranks; ++r) {
        c.peers.arena[r] = (ArenaHdr*)bases[r];
        c.peers.halo[r] = (double*)((char*)bases[r] + P2P_HDR_BYTES);
    }
    return comm_arena_finish(c);
}
inline int comm_arena_finish(Comm& c) {
    if (c.n_nbr > P2P_MAXR) { comm_err_ref() = "too many neighbours"; return 1; }
    c.halo.n_nbr = c.n_nbr;
    for (int k = 0; k < c.n_nbr; ++k) { c.halo.nbr_rank[k] = c.nbr_rank[k]; c.halo.ptr[k] = c.shared_ptr[k]; }
    c.halo.ptr[c.n_nbr] = c.n_nbr > 0 ? c.shared_ptr[c.n_nbr] : 0;
    c.halo.nodes = c.d_shared;
    for (int k = 0; k < c.n_nbr; ++k)
        if (3LL * (c.shared_ptr[k + 1] - c.shared_ptr[k]) > c.peers.cap) { comm_err_ref() = "halo capacity too small"; return 1; }
    {
        // spin time-out of the peer-memory collectives (default ~2 s; a debugger or profiler may need more)
        c.peers.timeout = P2P_TIMEOUT_CYCLES;
        if (const char* e = getenv("EN234_P2P_TIMEOUT_MS")) {
            const double ms = atof(e);
            if (ms > 0.0) c.peers.timeout = (long long)(ms * 2.0e6);
        }
    }
    // device copy of the peer table inside the local arena (producer tails reach it through one pointer)
    if (cudaMemcpy(&((ArenaHdr*)c.arena)->peers, &c.peers, sizeof(PeerSet), cudaMemcpyHostToDevice) != cudaSuccess) {
        comm_err_ref() = "cannot store the peer table";
        return 1;
    }
    // fused collectives: possible when the interface nodes are exactly the first local rows
    c.fused = false;
    c.F = FusedComm{};
    {
        // The interface tables are always built (rank-ordered halo sums use them).  Whether the PCG loop folds its
        // collectives into the kernels is c.fused_loop: measured on 2 and 8 B200 the folded forms were 4-13 % SLOWER than
        // single-CTA collective kernels between the compute kernels, so it is opt-in
        // (EN234_P2P_FUSED=1); ranks that are threads of one process use it too when asked.
        const char* e = getenv("EN234_P2P_FUSED");
        c.fused_loop = e && atoi(e) == 1;
        const bool want = true;
        if (want && (c.n_nbr == 0 || c.n_interface > 0)) {
            const int nif = c.n_interface;
            std::vector<int> cnt(nif + 1, 0);
            for (int k = 0; k < c.n_nbr; ++k)
                for (int q = c.shared_ptr[k]; q < c.shared_ptr[k + 1]; ++q) cnt[c.h_shared[q] + 1]++;
            for (int i = 0; i < nif; ++i) cnt[i + 1] += cnt[i];
            std::vector<int> fill(cnt.begin(), cnt.end() - 1), ifn(cnt[nif]), ifp(cnt[nif]);
            std::vector<unsigned char> before(nif > 0 ? nif : 1, 0);
            for (int k = 0; k < c.n_nbr; ++k)                     // neighbours are listed in ascending rank
                for (int q = c.shared_ptr[k]; q < c.shared_ptr[k + 1]; ++q) {
                    const int i = c.h_shared[q];
                    ifn[fill[i]] = k; ifp[fill[i]] = q - c.shared_ptr[k]; fill[i]++;
                    if (c.nbr_rank[k] < c.rank) before[i]++;
                }
            bool ascending = true;
            for (int k = 1; k < c.n_nbr; ++k) ascending = ascending && c.nbr_rank[k] > c.nbr_rank[k - 1];
            auto up = [](void** d, const void* h, size_t bytes) {
                if (cudaMalloc(d, bytes > 0 ? bytes : 4) != cudaSuccess) return 1;
                return bytes == 0 || cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 1;
            };
            if (ascending) {
                if (up((void**)&c.d_if_ptr, cnt.data(), cnt.size() * sizeof(int)) || up((void**)&c.d_if_nbr, ifn.data(), ifn.size() * sizeof(int)) ||
                    up((void**)&c.d_if_pos, ifp.data(), ifp.size() * sizeof(int)) || up((void**)&c.d_if_before, before.data(), before.size())) {
                    comm_err_ref() = "cudaMalloc failed for the interface tables";
                    return 1;
                }
                c.F.active = 1;
                c.F.P = c.peers;
                c.F.n_nbr = c.n_nbr;
                for (int k = 0; k < c.n_nbr; ++k) c.F.nbr_rank[k] = c.nbr_rank[k];
                c.F.n_if = nif;
                c.F.if_ptr = c.d_if_ptr; c.F.if_nbr = c.d_if_nbr; c.F.if_pos = c.d_if_pos; c.F.if_before = c.d_if_before;
                c.fused = true;
            }
        }
    }
    if (cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.ev_if, cudaEventDisableTiming) != cudaSuccess) {
        comm_err_ref() = "cannot create the side stream";
        return 1;
    }
    c.p2p = true;
    c.launches_per_halo = c.n_nbr > 0 ? (c.fused ? 2 : 1 + c.n_nbr) : 0;      // push + (one wait-and-total | one wait-and-add per neighbour)
    return 0;
}

inline int comm_p2p_error(Comm& c) {
    if (!c.p2p) return 0;
    int e = 0;
    cudaMemcpy(&e, &((ArenaHdr*)c.arena)->error, sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}

#define NCK(call)                                                                                    \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) { comm_err_ref() = std::string(#call) + ": " + api->GetErrorString(r_); return 1; } \
    } while (0)

inline int comm_unique_id(char* id128) {
    NcclApi* api = nccl_api();
    if (!api) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCK(api->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return 0;
}

inline int comm_init(Comm& c, int n_ranks, int rank, const char* id128, cudaStream_t) {
    c.n_ranks = n_ranks; c.rank = rank;
    if (n_ranks <= 1) { c.active = false; return 0; }
    NcclApi* api = nccl_api();
    if (!api) return 1;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCK(api->CommInitRank(&c.nccl, n_ranks, id, rank));
    c.active = true;
    return 0;
}

inline int comm_init_inproc(Comm& c, int n_ranks, int rank) {
    c.n_ranks = n_ranks; c.rank = rank;
    c.inproc = true;
    c.active = n_ranks > 1;
    return 0;
}

inline void comm_destroy(Comm& c) {
    if (c.nccl) { NcclApi* api = nccl_api(); if (api) api->CommDestroy(c.nccl); c.nccl = nullptr; }
    if (c.d_shared) cudaFree(c.d_shared);
    if (c.sendbuf) cudaFree(c.sendbuf);
    if (c.recvbuf) cudaFree(c.recvbuf);
    for (int r = 0; r < P2P_MAXR; ++r) if (c.peer_base[r]) { cudaIpcCloseMemHandle(c.peer_base[r]); c.peer_base[r] = nullptr; }
    if (c.arena) cudaFree(c.arena);
    if (c.d_scratch) cudaFree(c.d_scratch);
    c.d_scratch = nullptr;
    if (c.d_if_ptr) cudaFree(c.d_if_ptr);
    if (c.d_if_nbr) cudaFree(c.d_if_nbr);
    if (c.d_if_pos) cudaFree(c.d_if_pos);
    if (c.d_if_before) cudaFree(c.d_if_before);
    c.d_if_ptr = c.d_if_nbr = c.d_if_pos = nullptr; c.d_if_before = nullptr; c.fused = false; c.F = FusedComm{};
    if (c.side) cudaStreamDestroy(c.side);
    if (c.ev_fork) cudaEventDestroy(c.ev_fork);
    if (c.ev_join) cudaEventDestroy(c.ev_join);
    if (c.ev_if) cudaEventDestroy(c.ev_if);
    c.arena = nullptr; c.side = nullptr; c.ev_fork = c.ev_join = c.ev_if = nullptr; c.p2p = false;
    c.d_shared = nullptr; c.sendbuf = c.recvbuf = nullptr; c.active = false;
}

inline int comm_set_halo(Comm& c, int n_nbr, const int* nbr_rank, const int* shared_ptr, const int* shared_nodes) {
    c.n_nbr = n_nbr;
    c.nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
    c.shared_ptr.assign(shared_ptr, shared_ptr + n_nbr + 1);
    c.n_shared = n_nbr > 0 ? shared_ptr[n_nbr] : 0;
    c.h_shared.assign(shared_nodes, shared_nodes + c.n_shared);
    c.launches_per_halo = n_nbr > 0 ? 1 + n_nbr : 0;
    // interface rows are separable when the shared nodes are exactly the local nodes [0, n_interface)
    {
        int mx = -1;
        for (int i = 0; i < c.n_shared; ++i) mx = shared_nodes[i] > mx ? shared_nodes[i] : mx;
        std::vector<char> seen(mx + 1, 0);
        int distinct = 0;
        for (int i = 0; i < c.n_shared; ++i) if (!seen[shared_nodes[i]]) { seen[shared_nodes[i]] = 1; ++distinct; }
        c.n_interface = (distinct == mx + 1) ? mx + 1 : 0;
    }
    if (c.n_shared > 0) {
        if (cudaMalloc((void**)&c.d_shared, c.n_shared * sizeof(int)) != cudaSuccess ||
            cudaMalloc((void**)&c.sendbuf, 3 * (size_t)c.n_shared * sizeof(double)) != cudaSuccess ||
            cudaMalloc((void**)&c.recvbuf, 3 * (size_t)c.n_shared * sizeof(double)) != cudaSuccess) {
            comm_err_ref() = "cudaMalloc failed for halo buffers";
            return 1;
        }
        cudaMemcpy(c.d_shared, shared_nodes, c.n_shared * sizeof(int), cudaMemcpyHostToDevice);
    }
    return 0;
}

__global__ void k_halo_pack(int n, const int* nodes, const double* vec, double* buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    buf[t] = vec[3 * (size_t)nodes[t / 3] + t % 3];
}
__global__ void k_halo_add(int n, const int* nodes, const double* buf, double* vec) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    vec[3 * (size_t)nodes[t / 3] + t % 3] += buf[t];
}

// vec[shared nodes] += sum over neighbours of their partial values.  Both sides add the same two numbers
// (a+b == b+a), so the copies of an interface value stay bitwise identical on 1-D slab partitions.
// peer-memory halo sum on stream s (used where NCCL is not available: ranks that are threads of one process)
inline int comm_halo_sum_p2p(Comm& c, double* vec, cudaStream_t s) {
    if (c.n_nbr == 0) return 0;
    int mx = 0;
    for (int k = 0; k < c.n_nbr; ++k) mx = std::max(mx, 3 * (c.shared_ptr[k + 1] - c.shared_ptr[k]));
    const int nblk = std::max(1, std::min(32, (mx + 255) / 256));
    k_halo_push<<<dim3(nblk, c.n_nbr), 256, 0, s>>>(c.peers, c.halo, vec, nullptr, 0);
    if (c.fused) {
        k_halo_wait_total<<<std::max(1, std::min(64, (3 * c.n_interface + 255) / 256)), 256, 0, s>>>(c.F, vec);
    } else {
        for (int k = 0; k < c.n_nbr; ++k)
            k_halo_wait_add<<<nblk, 256, 0, s>>>(c.peers, c.halo, k, k == c.n_nbr - 1 ? 1 : 0, vec, nullptr, 0);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : (comm_err_ref() = "peer-memory halo sum: launch failed", 1);
}

inline int comm_halo_sum(Comm& c, double* vec, cudaStream_t s) {
    if (!c.active || c.n_shared == 0) return 0;
    // peer-memory arenas connected: every halo sum goes through them (assembly, boundary conditions, BiCGSTAB, the explicit
    // dynamic step), not only the PCG loop — two small kernels instead of pack + grouped send/recv + one add per neighbour
    if (c.p2p) return comm_halo_sum_p2p(c, vec, s);
    if (!c.nccl) { comm_err_ref() = "halo sum: no communicator"; return 1; }
    NcclApi* api = nccl_api();
    if (!api) return 1;
    k_halo_pack<<<(3 * c.n_shared + 255) / 256, 256, 0, s>>>(c.n_shared, c.d_shared, vec, c.sendbuf);
    NCK(api->GroupStart());
    for (int k = 0; k < c.n_nbr; ++k) {
        const size_t off = 3 * (size_t)c.shared_ptr[k], cnt = 3 * (size_t)(c.shared_ptr[k + 1] - c.shared_ptr[k]);
        NCK(api->Send(c.sendbuf + off, cnt, ncclDouble, c.nbr_rank[k], c.nccl, s));
        NCK(api->Recv(c.recvbuf + off, cnt, ncclDouble, c.nbr_rank[k], c.nccl, s));
    }
    NCK(api->GroupEnd());
    for (int k = 0; k < c.n_nbr; ++k) {           // one launch per neighbour: a node shared with two neighbours is
        const int off = c.shared_ptr[k], cnt = c.shared_ptr[k + 1] - off;   // updated in neighbour order, race-free
        k_halo_add<<<(3 * cnt + 255) / 256, 256, 0, s>>>(cnt, c.d_shared + off, c.recvbuf + 3 * (size_t)off, vec);
    }
    return 0;
}

inline int comm_allreduce_sum(Comm& c, double* dev, int n, cudaStream_t s) {
    if (!c.active) return 0;
    if (!c.nccl) {
        if (!c.p2p || n > 2) { comm_err_ref() = "all-reduce: no communicator"; return 1; }
        k_p2p_allreduce_vals<<<1, 32, 0, s>>>(c.peers, dev, n);
        return 0;
    }
    NcclApi* api = nccl_api();
    if (!api) return 1;
    NCK(api->AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, c.nccl, s));
    return 0;
}

// host int max across ranks (fail flags); uses a tiny device buffer
inline int comm_host_max(Comm& c, int* v, cudaStream_t s = nullptr) {
    if (!c.active) return 0;
    if (!c.nccl) {
        if (!c.p2p) { comm_err_ref() = "all-reduce: no communicator"; return 1; }
        const double x = *v != 0 ? 1.0 : 0.0;                       // sum of 0 / 1 flags: non-zero iff any rank raised one
        double y = 0.0;
        if (cudaMemcpyAsync(c.d_scratch, &x, sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
        k_p2p_allreduce_vals<<<1, 32, 0, s>>>(c.peers, c.d_scratch, 1);
        if (cudaMemcpyAsync(&y, c.d_scratch, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) return 1;
        if (cudaStreamSynchronize(s) != cudaSuccess) { comm_err_ref() = "all-reduce failed"; return 1; }
        *v = y != 0.0 ? 1 : 0;
        return 0;
    }
    NcclApi* api = nccl_api();
    if (!api) return 1;
    int* d = nullptr;
    cudaMalloc((void**)&d, sizeof(int));
    cudaMemcpy(d, v, sizeof(int), cudaMemcpyHostToDevice);
    NCK(api->AllReduce(d, d, 1, ncclInt, ncclMax, c.nccl, 0));
    cudaMemcpy(v, d, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}

// ---- small kernels used only by the partitioned path -----------------------------------------------------------
__global__ void __launch_bounds__(VEC_THREADS) k_negate_norm(int n, double* rhs, const double* felem, double* rforce,
                                                              const uint8_t* owned, double* partials) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        double r = rhs[d];
        if (felem) { r += felem[d]; rhs[d] = r; }
        rforce[d] = -r;
        if (owned[d / 3]) s += r * r;
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
__global__ void __launch_bounds__(VEC_THREADS) k_reduce2_to(const double* pa, const double* pb, int n, double* oa, double* ob) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    const double a = reduce_partials<VEC_THREADS>(pa, n, red);
    const double b = reduce_partials<VEC_THREADS>(pb, n, red);
    if (threadIdx.x == 0) { *oa = a; *ob = b; }
}
__global__ void k_pack2(double* a, const double* b) { a[1] = b[0]; }
__global__ void k_unpack2(const double* a, double* b) { b[0] = a[1]; }

// finalize the BC application after the halo sums of the correction and diagonal vectors
__global__ void __launch_bounds__(VEC_THREADS) k_bc_finalize(int n, double* rhs, const double* fext, const uint8_t* cflag,
                                                              const double* cval, const double* corr, const double* diag,
                                                              double* minv, const uint8_t* owned, double* partials) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        double r;
        if (cflag[d]) r = cval[d];
        else { r = rhs[d]; if (fext) r += fext[d]; r -= corr[d]; }
        rhs[d] = r;
        minv[d] = 1.0 / diag[d];
        if (owned == nullptr || owned[d / 3]) s += r * r;
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// BC application on a partition: local pass, halo sums of the additive vectors, finalisation
inline int bc_partitioned(Comm& c, BcArgs B, int grid_rows, int grid_vec, int ndof, const uint8_t* owned, double* corr,
                          double* diag, const double* assembled_diag, double* partials, cudaStream_t s, long long* launches) {
    B.corr_out = corr; B.diag_out = diag; B.owned = owned;
    if (B.rows) {
        // rows outside the subset: no correction, assembled diagonal
        if (cudaMemsetAsync(corr, 0, (size_t)ndof * sizeof(double), s) != cudaSuccess) return 1;
        if (cudaMemcpyAsync(diag, assembled_diag, (size_t)ndof * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return 1;
    }
    if (!B.rows || B.n_rows > 0) k_apply_bc<<<grid_rows, THREADS, 0, s>>>(B);
    if (comm_halo_sum(c, corr, s)) return 1;
    if (comm_halo_sum(c, diag, s)) return 1;
    k_bc_finalize<<<grid_vec, VEC_THREADS, 0, s>>>(ndof, B.rhs, B.fext, B.cflag, B.cval, corr, diag, B.minv, owned, partials);
    *launches += 2 + 2 * c.launches_per_halo;
    return 0;
}

}  // namespace en234k
