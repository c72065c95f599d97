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
constexpr int P2P_MAXCTA = 1024;          // CTAs of a consumer kernel that can be served by the local broadcast slots
constexpr long long P2P_TIMEOUT_CYCLES = 4000000000LL;   // default ~2 s at 2 GHz (EN234_P2P_TIMEOUT_MS overrides)

struct BcastSlot { double a, b; unsigned long long seq, pad; };     // one 32-byte sector per consumer CTA
struct ArenaHdr;

struct PeerSet {
    int n_ranks, me;
    ArenaHdr* arena[P2P_MAXR];     // arena[me] is local
    double* halo[P2P_MAXR];        // halo receive area of each rank: [source rank][parity][cap]
    long long cap;                 // doubles per (source, parity)
    long long timeout;             // spin time-out in SM clock cycles
};

struct ArenaHdr {
    unsigned long long ar_flag[2][P2P_MAXR];   // all-reduce: sequence number per (parity, source rank)
    double ar_val[2][P2P_MAXR][4];
    unsigned long long halo_flag[P2P_MAXR];    // halo: sequence number of the last completed push per source rank
    unsigned long long ar_seq, halo_seq;       // local operation counters
    unsigned int ticket[P2P_MAXR];             // per-neighbour CTA tickets of k_halo_push
    unsigned int ticket_wait;
    int error;
    unsigned int tk_spmv, tk_upd;              // last-CTA tickets of the fused PCG kernels (k_spmv tail, k_cg_update tail)
    unsigned int arrive;                       // warps of the fused k_spmv that have finished their interface rows
    int pad[3];
    unsigned long long ll[2][P2P_MAXR][4];     // all-reduce packets (k_p2p_allreduce): 4 words of (sequence << 32 | half a double)
    BcastSlot bc[2][P2P_MAXCTA];               // local re-broadcast of an all-reduced pair to the CTAs of the consumer kernel
    PeerSet peers;                             // device copy of the peer table (producer tails reach it through one pointer)
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
// A time-out (a peer stopped responding) raises arena->error AND the solver's stop / fail flags, so every later kernel of
// the captured iterations returns at once instead of spinning again; the communicator must be rebuilt afterwards (the
// sequence counters of the ranks no longer agree).
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long seq, int* error,
                                           long long timeout = P2P_TIMEOUT_CYCLES, CgScalars* S = nullptr) {
    const long long t0 = clock64();
    while (ld_volatile_u64(flag) < seq) {
        if (clock64() - t0 > timeout) {
            *(volatile int*)error = 1;
            if (S) { *(volatile int*)&S->fail = 1; *(volatile int*)&S->stop = 1; }
            return false;
        }
        __nanosleep(20);
    }
    return true;
}

// single CTA.  outA/outB receive the global sums; partB may be null.
// The pair travels as four 8-byte packets, each carrying its own validity tag — (sequence number << 32) | half of a double,
// one aligned 8-byte store each — so no fence separates "data" from "flag" (a system-scope fence between a remote store and
// its flag costs an NVLink round trip): the receiver polls the four words of every source until all carry this operation's
// sequence number.  Slots alternate with the sequence parity; a slot's previous content carries sequence - 2.
// Measured on 2 B200 (tests/mgpu_check.py --latency-only): 8.0 us per call with value / fence / flag, see DESIGN.md section 5.
__device__ __forceinline__ unsigned long long ll_pack(unsigned long long seq, unsigned int half) {
    return (seq << 32) | (unsigned long long)half;
}
__global__ void __launch_bounds__(VEC_THREADS) k_p2p_allreduce(PeerSet P, const double* partA, int nA, const double* partB,
                                                               int nB, double* outA, double* outB, const CgScalars* S,
                                                               int check_stop) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    __shared__ double red2[VEC_THREADS / 32 + 1];
    __shared__ double got[P2P_MAXR][2];
    pdl_wait();
    pdl_trigger();
    if (check_stop && S->stop) return;
    ArenaHdr* mine = P.arena[P.me];
    // both partial arrays in one pass; per value the same order as reduce_partials (strided sums, shuffle tree, warps in order)
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nA; i += VEC_THREADS) a += partA[i];
    if (partB) for (int i = threadIdx.x; i < nB; i += VEC_THREADS) b += partB[i];
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = a; red2[threadIdx.x >> 5] = b; }
    __syncthreads();
    const unsigned long long seq = ld_volatile_u64(&mine->ar_seq) + 1ULL;
    const unsigned long long tag = seq & 0xffffffffULL;
    const int par = (int)(seq & 1ULL);
    const int r = threadIdx.x;
    if (r < P.n_ranks) {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int w = 0; w < VEC_THREADS / 32; ++w) { sa += red[w]; sb += red2[w]; }
        const unsigned long long ua = (unsigned long long)__double_as_longlong(sa), ub = (unsigned long long)__double_as_longlong(sb);
        volatile unsigned long long* dst = P.arena[r]->ll[par][P.me];
        dst[0] = ll_pack(tag, (unsigned int)ua);
        dst[1] = ll_pack(tag, (unsigned int)(ua >> 32));
        dst[2] = ll_pack(tag, (unsigned int)ub);
        dst[3] = ll_pack(tag, (unsigned int)(ub >> 32));
        // wait for source rank r's four packets
        const volatile unsigned long long* src = mine->ll[par][r];
        const long long t0 = clock64();
        unsigned long long w0, w1, w2, w3;
        bool ok = true;
        for (;;) {
            w0 = src[0]; w1 = src[1]; w2 = src[2]; w3 = src[3];
            if ((w0 >> 32) == tag && (w1 >> 32) == tag && (w2 >> 32) == tag && (w3 >> 32) == tag) break;
            if (clock64() - t0 > P.timeout) {
                *(volatile int*)&mine->error = 1;
                if (check_stop) { CgScalars* Sw = const_cast<CgScalars*>(S); *(volatile int*)&Sw->fail = 1; *(volatile int*)&Sw->stop = 1; }
                ok = false;
                break;
            }
        }
        got[r][0] = ok ? __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32))) : 0.0;
        got[r][1] = ok ? __longlong_as_double((long long)((w2 & 0xffffffffULL) | (w3 << 32))) : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int q = 0; q < P.n_ranks; ++q) { sa += got[q][0]; sb += got[q][1]; }      // fixed rank order => identical sums on all ranks
        *outA = sa;
        if (outB) *outB = sb;
        mine->ar_seq = seq;
    }
}

// the first form of the all-reduce kernel (value, system fence, flag): kept for A/B (EN234_P2P_LL=0)
__global__ void __launch_bounds__(VEC_THREADS) k_p2p_allreduce_fence(PeerSet P, const double* partA, int nA, const double* partB,
                                                               int nB, double* outA, double* outB, const CgScalars* S,
                                                               int check_stop) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    __shared__ unsigned long long s_seq;
    pdl_wait();
    pdl_trigger();
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
        spin_until(&mine->ar_flag[par][r], seq, &mine->error, P.timeout, check_stop ? const_cast<CgScalars*>(S) : nullptr);
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
    if (threadIdx.x == 0) spin_until(&mine->halo_flag[nb], seq, &mine->error, P.timeout, check_stop ? const_cast<CgScalars*>(S) : nullptr);
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

// ---- collectives folded into the PCG kernels (no kernel of their own) ---------------------------------------------------
// The producer of a dot product (k_spmv for p.Ap, k_cg_update for r.r / r.z) ends with a last-CTA ticket: that CTA reduces
// the kernel's per-CTA partials in the fixed order and stores the rank's sum into every rank's arena slot, then the flag.
// The consumer kernel (k_cg_update / k_cg_pupdate) issues its first vector loads, then every CTA spins on the LOCAL
// arena flags and adds the slots in rank order: bitwise identical sums on all ranks, no single-CTA all-reduce launch on
// the critical path.  Sequence numbers: arena->ar_seq counts completed all-reduces; producers use ar_seq + 1, the
// consumer's last CTA (ticket) advances it after every CTA has read it.  Two parity slots suffice: a rank can only push
// sequence s + 2 after it consumed s + 1, which its peers pushed after they had consumed s.
// Halo sums: the warps of k_spmv store the A p values of interface rows straight into the neighbours' receive areas and
// count themselves off once their interface rows are done (interface rows are the first local rows); the last warp
// raises the neighbours' flags.  The consumer k_cg_update adds what arrived to its own partial value in ascending RANK
// order (its own contribution inserted at its rank), so every copy of an interface value is the bitwise identical sum on
// every rank that holds it, also for nodes shared by more than two ranks.
struct FusedComm {
    int active;                       // 0: one GPU, NCCL path or the unfused peer-memory path
    int halo_in_spmv;                 // 1: k_spmv pushes the interface rows itself (assembled operator)
    int pAp_pushed;                   // 1: the operator kernel's tail has pushed the rank's p.Ap (assembled operator)
    PeerSet P;
    int n_nbr;
    int nbr_rank[P2P_MAXR];
    int n_if;                         // interface nodes are the local rows [0, n_if)
    const int* if_ptr;                // n_if + 1: entries of interface node i
    const int* if_nbr;                // neighbour slot k of the entry; a node's entries are in ascending neighbour rank
    const int* if_pos;                // position of the node in the list shared with that neighbour
    const unsigned char* if_before;   // per interface node: number of its entries whose rank is below this rank
};

__device__ __forceinline__ double* halo_send_area(const FusedComm& F, int k, int par) {
    return F.P.halo[F.nbr_rank[k]] + ((long long)F.P.me * 2 + par) * F.P.cap;
}
__device__ __forceinline__ const double* halo_recv_area(const FusedComm& F, int k, int par) {
    return F.P.halo[F.P.me] + ((long long)F.nbr_rank[k] * 2 + par) * F.P.cap;
}

// all threads of the CTA; a and b are CTA-uniform
__device__ __forceinline__ void p2p_push2(const PeerSet& P, unsigned long long seq, double a, double b) {
    const int par = (int)(seq & 1ULL);
    const int r = threadIdx.x;
    if (r < P.n_ranks) {
        ArenaHdr* dst = P.arena[r];
        *(volatile double*)&dst->ar_val[par][P.me][0] = a;
        *(volatile double*)&dst->ar_val[par][P.me][1] = b;
        __threadfence_system();
        *(volatile unsigned long long*)&dst->ar_flag[par][P.me] = seq;
    }
}

// Tail of a producer kernel (k_spmv, k_cg_update), all threads of every CTA: the CTA that finishes last — counted over
// `total` CTAs, which may belong to two concurrent launches — reduces the kernel's per-CTA partials in the fixed order and
// stores the rank's pair into every rank's arena.  The peers' stores cross NVLink while the consumer kernel is being
// launched.  `arena` is the local arena (its device copy of the peer table is used), `ticket` one of its counters.
template <int NT>
__device__ __forceinline__ void p2p_tail_push(ArenaHdr* arena, unsigned int* ticket, unsigned int total, const double* partA,
                                              const double* partB, int n, double* red, double* outA = nullptr,
                                              double* outB = nullptr, CgScalars* S = nullptr) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == total - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < n; i += NT) { a += __ldcg(partA + i); if (partB) b += __ldcg(partB + i); }
        a = block_sum<NT>(a, red);
        b = block_sum<NT>(b, red);
        if (threadIdx.x == 0) *ticket = 0;
        const unsigned long long seq = ld_volatile_u64(&arena->ar_seq) + 1ULL;
        p2p_push2(arena->peers, seq, a, b);
        if (outA) {
            // EN234_P2P_TAIL=2: the same CTA is also the consumer half (no collective kernel at all): wait for every rank's
            // pair, add in rank order, leave the sums where the next kernel reads its single partial
            const int par = (int)(seq & 1ULL);
            const int r = threadIdx.x;
            if (r < arena->peers.n_ranks) { spin_until(&arena->ar_flag[par][r], seq, &arena->error, arena->peers.timeout, S); __threadfence_system(); }
            __syncthreads();
            if (r == 0) {
                double sa = 0.0, sb = 0.0;
                for (int q = 0; q < arena->peers.n_ranks; ++q) {
                    sa += *(volatile double*)&arena->ar_val[par][q][0];
                    sb += *(volatile double*)&arena->ar_val[par][q][1];
                }
                *outA = sa;
                if (outB) *outB = sb;
                arena->ar_seq = seq;
            }
        }
    }
}

// All-reduce of the pair (sum of partA, sum of partB) inside the consumer kernel, all threads of every CTA (>= 64 threads).
// The rank's pair normally arrives from the producer's tail (pushed != 0); otherwise CTA 0 reduces the partials and pushes.
// Every CTA first looks ONCE at the peers' flags (and, when it handles interface DOFs, at the neighbours' halo flags): if
// everything has arrived — the usual case, the stores crossed NVLink during the kernel launch — it adds the pairs in rank
// order itself.  Otherwise it waits for CTA 0, the rank's agent, which spins on the flags and re-broadcasts the result
// through one 32-byte slot per CTA in LOCAL memory: no flag is polled by more than a handful of threads (hundreds of CTAs
// spinning on the one line a peer is trying to write cost 15-20 us per all-reduce on B200).
// sh: 2 doubles of shared memory, red: scratch of reduce_partials.
template <int NT>
__device__ __forceinline__ void p2p_allreduce_cta(const FusedComm& F, unsigned long long seq, unsigned long long hseq, CgScalars* S,
                                                  int pushed, bool need_halo, const double* partA, int nA, const double* partB,
                                                  int nB, double* red, double* sh, double& a, double& b) {
    ArenaHdr* mine = F.P.arena[F.P.me];
    const int par = (int)(seq & 1ULL);
    const int t = threadIdx.x;
    BcastSlot* bc = mine->bc[par];
    const bool agent = blockIdx.x == 0;
    if (agent && !pushed) {
        const double la = reduce_partials<NT>(partA, nA, red);
        const double lb = partB ? reduce_partials<NT>(partB, nB, red) : 0.0;
        p2p_push2(F.P, seq, la, lb);
    }
    // one look at the flags this CTA depends on
    int here = 1;
    if (t < F.P.n_ranks) here = ld_volatile_u64(&mine->ar_flag[par][t]) >= seq;
    else if ((need_halo || agent) && hseq != 0ULL && t >= 32 && t < 32 + F.n_nbr)
        here = ld_volatile_u64(&mine->halo_flag[F.nbr_rank[t - 32]]) >= hseq;
    if (here) __threadfence_system();                              // acquire (passed on by the barrier)
    const int all_here = __syncthreads_and(here);
    if (!all_here && agent) {
        if (t < F.P.n_ranks) { spin_until(&mine->ar_flag[par][t], seq, &mine->error, F.P.timeout, S); __threadfence_system(); }
        else if (hseq != 0ULL && t >= 32 && t < 32 + F.n_nbr) {
            spin_until(&mine->halo_flag[F.nbr_rank[t - 32]], hseq, &mine->error, F.P.timeout, S);
            __threadfence_system();
        }
        __syncthreads();
    }
    if (all_here || agent) {
        if (t == 0) {
            double sa = 0.0, sb = 0.0;
            for (int q = 0; q < F.P.n_ranks; ++q) {                    // fixed rank order => identical sums on all ranks
                sa += *(volatile double*)&mine->ar_val[par][q][0];
                sb += *(volatile double*)&mine->ar_val[par][q][1];
            }
            sh[0] = sa; sh[1] = sb;
        }
        __syncthreads();
        a = sh[0]; b = sh[1];
        if (agent) {                                                    // always: some CTA may have looked too early
            for (int j = 1 + t; j < (int)gridDim.x; j += NT) {
                *(volatile double*)&bc[j].a = a;
                *(volatile double*)&bc[j].b = b;
            }
            __threadfence();
            for (int j = 1 + t; j < (int)gridDim.x; j += NT) *(volatile unsigned long long*)&bc[j].seq = seq;
        }
    } else {
        if (t == 0) {
            spin_until(&bc[blockIdx.x].seq, seq, &mine->error, F.P.timeout, S);
            __threadfence();
            sh[0] = *(volatile double*)&bc[blockIdx.x].a;
            sh[1] = *(volatile double*)&bc[blockIdx.x].b;
        }
        __syncthreads();
        a = sh[0]; b = sh[1];
    }
}

// value of interface DOF d (= 3 * node + c, node < n_if): own partial value and the neighbours' in ascending rank order
__device__ __forceinline__ double halo_total(const FusedComm& F, int hpar, int d, double own) {
    const int i = d / 3, c = d - 3 * i;
    const int e0 = F.if_ptr[i], cnt = F.if_ptr[i + 1] - e0, before = F.if_before[i];
    double s = 0.0;
    for (int t = 0; t <= cnt; ++t) {
        double v;
        if (t == before) v = own;
        else {
            const int j = e0 + (t > before ? t - 1 : t);
            v = __ldcv(halo_recv_area(F, F.if_nbr[j], hpar) + 3 * (size_t)F.if_pos[j] + c);
        }
        s = t == 0 ? v : s + v;
    }
    return s;
}

// vec[interface DOFs] = own partial value + what the neighbours pushed, in ascending rank order (halo_total); waits for all
// neighbours' flags first; the last CTA advances the halo sequence counter.  Replaces the per-neighbour k_halo_wait_add
// launches where the interface tables exist.
__global__ void __launch_bounds__(256) k_halo_wait_total(FusedComm F, double* __restrict__ vec, CgScalars* S = nullptr,
                                                         int check_stop = 0) {
    if (check_stop && S->stop) return;
    ArenaHdr* mine = F.P.arena[F.P.me];
    const unsigned long long hseq = ld_volatile_u64(&mine->halo_seq) + 1ULL;
    const int hpar = (int)(hseq & 1ULL);
    if (threadIdx.x < F.n_nbr)
        spin_until(&mine->halo_flag[F.nbr_rank[threadIdx.x]], hseq, &mine->error, F.P.timeout, check_stop ? S : nullptr);
    __syncthreads();
    __threadfence_system();
    const int n3 = 3 * F.n_if;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n3; d += gridDim.x * blockDim.x) vec[d] = halo_total(F, hpar, d, vec[d]);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(&mine->ticket_wait, 1u);
        if (tk == gridDim.x - 1) { mine->ticket_wait = 0; mine->halo_seq = hseq; }
    }
}

// one warp: all-reduce (sum) of n <= 2 device scalars in place, through the arenas, added in rank order
__global__ void k_p2p_allreduce_vals(PeerSet P, double* vals, int n) {
    ArenaHdr* mine = P.arena[P.me];
    const unsigned long long seq = ld_volatile_u64(&mine->ar_seq) + 1ULL;
    const int par = (int)(seq & 1ULL);
    const double a = vals[0], b = n > 1 ? vals[1] : 0.0;
    __syncwarp();
    const int r = threadIdx.x;
    if (r < P.n_ranks) {
        ArenaHdr* dst = P.arena[r];
        *(volatile double*)&dst->ar_val[par][P.me][0] = a;
        *(volatile double*)&dst->ar_val[par][P.me][1] = b;
        __threadfence_system();
        *(volatile unsigned long long*)&dst->ar_flag[par][P.me] = seq;
        spin_until(&mine->ar_flag[par][r], seq, &mine->error, P.timeout);
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        __threadfence_system();
        double sa = 0.0, sb = 0.0;
        for (int q = 0; q < P.n_ranks; ++q) {
            sa += *(volatile double*)&mine->ar_val[par][q][0];
            sb += *(volatile double*)&mine->ar_val[par][q][1];
        }
        vals[0] = sa;
        if (n > 1) vals[1] = sb;
        mine->ar_seq = seq;
    }
}

// one warp: the consumer half of an all-reduce whose pairs were pushed by the producers' tails (p2p_tail_push): wait for
// all ranks' flags, add the pairs in rank order, store the sums, advance the counter
__global__ void k_p2p_wait(PeerSet P, double* outA, double* outB, CgScalars* S) {
    if (S->stop) return;
    ArenaHdr* mine = P.arena[P.me];
    const unsigned long long seq = ld_volatile_u64(&mine->ar_seq) + 1ULL;
    const int par = (int)(seq & 1ULL);
    const int r = threadIdx.x;
    if (r < P.n_ranks) { spin_until(&mine->ar_flag[par][r], seq, &mine->error, P.timeout, S); __threadfence_system(); }
    __syncwarp();
    if (r == 0) {
        double sa = 0.0, sb = 0.0;
        for (int q = 0; q < P.n_ranks; ++q) {
            sa += *(volatile double*)&mine->ar_val[par][q][0];
            sb += *(volatile double*)&mine->ar_val[par][q][1];
        }
        *outA = sa;
        if (outB) *outB = sb;
        mine->ar_seq = seq;
    }
}

// single CTA: reduce the partials of an operator kernel that has no tail of its own (matrix-free operators) and push
__global__ void __launch_bounds__(VEC_THREADS) k_p2p_push(PeerSet P, const double* partA, int nA, const CgScalars* S) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S->stop) return;
    const double a = reduce_partials<VEC_THREADS>(partA, nA, red);
    p2p_push2(P, ld_volatile_u64(&P.arena[P.me]->ar_seq) + 1ULL, a, 0.0);
}

}  // namespace en234k
