// Peer-memory collectives for the PCG loop of partitioned runs (one process per GPU, NVLink/NVSwitch).
// Every rank owns an "arena" in device memory whose CUDA-IPC handle is opened by all other ranks, so kernels store
// straight into their peers' buffers and signal with monotonically increasing sequence flags:
//   * halo sum:   k_halo_push writes this rank's interface values into each neighbour's receive area, then the flag;
//                 k_halo_wait_add waits for the neighbours' flags and adds what arrived.  Both run on a side stream
//                 and overlap the SpMV of the interior rows (interface rows are computed first).
//   * all-reduce: k_p2p_allreduce reduces the per-CTA partials of the producer kernel, stores the local sum into every
//                 rank's slot, waits for all slots and adds them in rank order — every rank obtains the bitwise
//                 identical sum.  One single-CTA kernel replaces reduce + ncclAllReduce (+ pack/unpack).
// Kernels that wait on one another run on DIFFERENT GPUs (never co-scheduled on one device).  Every spin has a
// time-out that raises arena->error instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>

#include "en234gpu_kernels.cuh"

namespace en234k {

constexpr int P2P_MAXR = 16;
constexpr long long P2P_TIMEOUT_CYCLES = 4000000000LL;   // ~2 s at 2 GHz

struct ArenaHdr {
    unsigned long long ar_flag[2][P2P_MAXR];   // all-reduce: sequence number per (parity, source rank)
    double ar_val[2][P2P_MAXR][4];
    unsigned long long halo_flag[P2P_MAXR];    // halo: sequence number of the last completed push per source rank
    unsigned long long ar_seq, halo_seq;       // local operation counters
    unsigned int ticket[P2P_MAXR];             // per-neighbour CTA tickets of k_halo_push
    unsigned int ticket_wait;
    int error;
    int pad[2];
};

struct PeerSet {
    int n_ranks, me;
    ArenaHdr* arena[P2P_MAXR];     // arena[me] is local
    double* halo[P2P_MAXR];        // halo receive area of each rank: [source rank][parity][cap]
    long long cap;                 // doubles per (source, parity)
};

struct HaloSet {
    int n_nbr;
    int nbr_rank[P2P_MAXR];
    int ptr[P2P_MAXR + 1];         // offsets into `nodes`
    const int* nodes;              // local node index per shared entry
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *(const volatile unsigned long long*)p;
}
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long seq, int* error) {
    const long long t0 = clock64();
    while (ld_volatile_u64(flag) < seq) {
        if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { *(volatile int*)error = 1; return false; }
        __nanosleep(20);
    }
    return true;
}

// single CTA.  outA/outB receive the global sums; partB may be null.
__global__ void __launch_bounds__(VEC_THREADS) k_p2p_allreduce(PeerSet P, const double* partA, int nA, const double* partB,
                                                               int nB, double* outA, double* outB, const CgScalars* S,
                                                               int check_stop) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    __shared__ unsigned long long s_seq;
    if (check_stop && S->stop) return;
    ArenaHdr* mine = P.arena[P.me];
    const double a = reduce_partials<VEC_THREADS>(partA, nA, red);
    const double b = partB ? reduce_partials<VEC_THREADS>(partB, nB, red) : 0.0;
    if (threadIdx.x == 0) s_seq = mine->ar_seq + 1;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ULL);
    const int r = threadIdx.x;
    if (r < P.n_ranks) {
        ArenaHdr* dst = P.arena[r];
        *(volatile double*)&dst->ar_val[par][P.me][0] = a;
        *(volatile double*)&dst->ar_val[par][P.me][1] = b;
        __threadfence_system();
        *(volatile unsigned long long*)&dst->ar_flag[par][P.me] = seq;
        spin_until(&mine->ar_flag[par][r], seq, &mine->error);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        double sa = 0.0, sb = 0.0;
        for (int q = 0; q < P.n_ranks; ++q) {                      // fixed rank order => identical sums on all ranks
            sa += *(volatile double*)&mine->ar_val[par][q][0];
            sb += *(volatile double*)&mine->ar_val[par][q][1];
        }
        *outA = sa;
        if (outB) *outB = sb;
        mine->ar_seq = seq;
    }
}

// grid (nblk, n_nbr): copy this rank's interface values into neighbour blockIdx.y's receive area, then flag.
__global__ void __launch_bounds__(256) k_halo_push(PeerSet P, HaloSet H, const double* __restrict__ vec, const CgScalars* S,
                                                   int check_stop) {
    if (check_stop && S->stop) return;
    ArenaHdr* mine = P.arena[P.me];
    const int k = blockIdx.y;
    const int nb = H.nbr_rank[k];
    const unsigned long long seq = mine->halo_seq + 1;
    const int par = (int)(seq & 1ULL);
    const int off = H.ptr[k], cnt = 3 * (H.ptr[k + 1] - H.ptr[k]);
    double* dst = P.halo[nb] + ((long long)P.me * 2 + par) * P.cap;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += gridDim.x * blockDim.x)
        dst[t] = vec[3 * (size_t)H.nodes[off + t / 3] + t % 3];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&mine->ticket[k], 1u);
        if (tk == gridDim.x - 1) {
            mine->ticket[k] = 0;
            __threadfence_system();
            *(volatile unsigned long long*)&P.arena[nb]->halo_flag[P.me] = seq;
        }
    }
}

// vec[nodes shared with neighbour k] += what neighbour k pushed.  One launch per neighbour, in neighbour order, so a
// node shared with two neighbours is updated by one pass after the other (race-free, fixed order); the last launch
// advances the halo sequence counter.
__global__ void __launch_bounds__(256) k_halo_wait_add(PeerSet P, HaloSet H, int k, int last, double* __restrict__ vec,
                                                       const CgScalars* S, int check_stop) {
    if (check_stop && S->stop) return;
    ArenaHdr* mine = P.arena[P.me];
    const unsigned long long seq = mine->halo_seq + 1;
    const int par = (int)(seq & 1ULL);
    const int nb = H.nbr_rank[k];
    if (threadIdx.x == 0) spin_until(&mine->halo_flag[nb], seq, &mine->error);
    __syncthreads();
    __threadfence_system();
    const int off = H.ptr[k], cnt = 3 * (H.ptr[k + 1] - H.ptr[k]);
    const double* src = P.halo[P.me] + ((long long)nb * 2 + par) * P.cap;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += gridDim.x * blockDim.x) {
        const size_t d = 3 * (size_t)H.nodes[off + t / 3] + t % 3;
        vec[d] += __ldcv(src + t);
    }
    if (last) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int tk = atomicAdd(&mine->ticket_wait, 1u);
            if (tk == gridDim.x - 1) { mine->ticket_wait = 0; mine->halo_seq = seq; }
        }
    }
}

}  // namespace en234k
