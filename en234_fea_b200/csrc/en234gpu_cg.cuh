// Jacobi-preconditioned conjugate gradient kernels: the recurrences of cg_iteration_loop
// (src/solver_conjugate_gradient.f90:776-912) with the three global dot products fused into the kernels that
// produce their operands.  One PCG iteration = k_spmv -> k_update -> k_pupdate; all scalars live in device
// memory (CgScalars), per-CTA partial sums are reduced redundantly by every consumer CTA in a fixed order, so the
// iteration has no host round trip, no atomics and is bitwise run-to-run deterministic.
#pragma once
#include "en234gpu_kernels.cuh"
#include "en234gpu_p2p.cuh"

namespace en234k {

struct SpmvArgs {
    int row0 = 0;           // first block row of this launch (interface rows are launched first on partitioned runs)
    int N;                  // one past the last block row of this launch
    const int* browptr;
    const int* bcol;
    const double* val;
    const double* x;        // p
    double* y;              // A p
    double* partials;       // per-CTA sum of p.(A p) over the CTA's rows (all local rows, see DESIGN.md multi-GPU)
    const CgScalars* S;
    int check_stop;
    // partitioned runs with folded collectives: the CTA that finishes last among tail_total CTAs (of the interface and the
    // interior launch together) reduces tail_partials[0 .. tail_total) and pushes the rank's p.Ap (p2p_tail_push)
    ArenaHdr* tail = nullptr;
    const double* tail_partials = nullptr;
    int tail_total = 0;
    double* tail_out = nullptr;    // EN234_P2P_TAIL=2: the last CTA also waits for the peers' sums and stores the total here
};

// reduce_partials for partials written by OTHER CTAs of the same launch (last-CTA tails): loads bypass L1
template <int NT>
__device__ __forceinline__ double reduce_partials_cg(const double* part, int n, double* sm) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += NT) s += __ldcg(part + i);
    return block_sum<NT>(s, sm);
}

// y = A x for the 3x3-block CSR matrix stored per block row as [r][block][c]; one warp per block row.
// Lane q < 3*nb owns scalar column q of the block row: it gathers x_q once and reuses it for the three scalar
// rows, whose values are read as three fully coalesced streams (streaming loads: the matrix is read once per
// iteration and must not evict the vectors from L2).
// The index loads are software-pipelined across the warp's rows (row pointers two rows ahead, block columns one row
// ahead), so every load issued for a row is independent of the others: one memory latency per row instead of the
// pointer -> column -> x chain.  Rows are interleaved over all warps of the (single-wave, persistent) grid so that the
// rows in flight at any time are neighbours and their x entries hit in L2/L1.
// WIDE = false: no block row has more than 32 blocks (every hex8 mesh with at most 8 elements per node), the tail loop
// over further blocks is compiled out.
// FUSED (partitioned runs, peer-memory arenas; launched over the interface rows [0, F.n_if) only, the interior rows go
// through the lean variant so that the hot loop carries nothing extra): every row also stores its result into the
// neighbours' halo receive areas, every warp counts itself off when its rows are done and the last one raises the
// neighbours' flags (see en234gpu_p2p.cuh).
template <int MINB, bool WIDE = true, bool FUSED = false>
__global__ void __launch_bounds__(THREADS, MINB) k_spmv(SpmvArgs A, FusedComm F = FusedComm{}) {
    __shared__ double red[WARPS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ArenaHdr* mine = nullptr;
    unsigned long long hseq = 0ULL;
    bool arrived = true;
    if (FUSED) {
        mine = F.P.arena[F.P.me];
        hseq = ld_volatile_u64(&mine->halo_seq) + 1ULL;
        arrived = F.halo_in_spmv == 0;
    }
    auto arrive = [&]() {
        // this warp's interface rows are stored (locally and in the neighbours' arenas): count it off; the last warp
        // of the grid raises the flags
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            const unsigned t = atomicAdd(&mine->arrive, 1u);
            if (t == gridDim.x * WARPS - 1) {
                mine->arrive = 0;
                __threadfence_system();
                for (int k = 0; k < F.n_nbr; ++k)
                    *(volatile unsigned long long*)&F.P.arena[F.nbr_rank[k]]->halo_flag[F.P.me] = hseq;
            }
        }
        arrived = true;
    };
    const int W = gridDim.x * WARPS, N = A.N;          // rows [A.row0, A.N) are processed by this launch
    const int d0 = lane / 3, m0 = lane % 3;                  // q = lane      -> block d0, component m0
    const int d1 = (lane + 32) / 3, m1 = (lane + 32) % 3;    // q = lane + 32
    const int d2 = (lane + 64) / 3, m2 = (lane + 64) % 3;    // q = lane + 64
    double dot = 0.0;
    int i = A.row0 + blockIdx.x * WARPS + warp;
    int rb = 0, re = 0, rbn = 0, ren = 0, c0 = 0, c1 = 0, c2 = 0;
    if (i < N) { rb = __ldg(A.browptr + i); re = __ldg(A.browptr + i + 1); }
    if (i + W < N) { rbn = __ldg(A.browptr + i + W); ren = __ldg(A.browptr + i + W + 1); }
    {
        const int nb = re - rb;
        if (d0 < nb) c0 = __ldg(A.bcol + rb + d0);
        if (d1 < nb) c1 = __ldg(A.bcol + rb + d1);
        if (d2 < nb) c2 = __ldg(A.bcol + rb + d2);
    }
    // the index loads above touch nothing the previous kernel writes: with a programmatic launch they overlap its tail
    pdl_wait();
    pdl_trigger();
    if (A.check_stop && A.S->stop) return;
    for (; i < N; i += W) {
        if (FUSED && !arrived && i >= F.n_if) arrive();
        const int nb = re - rb, L = 3 * nb;
        const double* v = A.val + 9 * (size_t)rb;
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
        // ---- loads of this row: 3 x-gathers + 9 streaming value loads per lane, all independent
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, a00 = 0.0, a01 = 0.0, a02 = 0.0, a10 = 0.0, a11 = 0.0, a12 = 0.0, a20 = 0.0,
               a21 = 0.0, a22 = 0.0;
        if (d0 < nb) { x0 = __ldg(A.x + 3 * c0 + m0); a00 = __ldcs(v + lane); a10 = __ldcs(v + L + lane); a20 = __ldcs(v + 2 * L + lane); }
        if (d1 < nb) { x1 = __ldg(A.x + 3 * c1 + m1); a01 = __ldcs(v + lane + 32); a11 = __ldcs(v + L + lane + 32); a21 = __ldcs(v + 2 * L + lane + 32); }
        if (d2 < nb) { x2 = __ldg(A.x + 3 * c2 + m2); a02 = __ldcs(v + lane + 64); a12 = __ldcs(v + L + lane + 64); a22 = __ldcs(v + 2 * L + lane + 64); }
        // ---- prefetch: columns of the next row, pointers of the row after it
        const int nbn = ren - rbn;
        int cn0 = 0, cn1 = 0, cn2 = 0, rb2 = 0, re2 = 0;
        if (d0 < nbn) cn0 = __ldg(A.bcol + rbn + d0);
        if (d1 < nbn) cn1 = __ldg(A.bcol + rbn + d1);
        if (d2 < nbn) cn2 = __ldg(A.bcol + rbn + d2);
        if (i + 2 * W < N) { rb2 = __ldg(A.browptr + i + 2 * W); re2 = __ldg(A.browptr + i + 2 * W + 1); }
        y0 = a00 * x0 + a01 * x1 + a02 * x2;
        y1 = a10 * x0 + a11 * x1 + a12 * x2;
        y2 = a20 * x0 + a21 * x1 + a22 * x2;
        if (WIDE)
        for (int q = lane + 96; q < L; q += 32) {            // rows with more than 32 blocks (unstructured meshes)
            const int j = 3 * __ldg(A.bcol + rb + q / 3) + q % 3;
            const double xv = __ldg(A.x + j);
            y0 += __ldcs(v + q) * xv;
            y1 += __ldcs(v + L + q) * xv;
            y2 += __ldcs(v + 2 * L + q) * xv;
        }
        const double yv = warp_sum3(y0, y1, y2, lane);          // y_r in lanes 8r..8r+7
        if ((lane & 7) == 0 && lane < 24) {
            const int r = lane >> 3;
            A.y[3 * (size_t)i + r] = yv;
            dot += yv * A.x[3 * (size_t)i + r];
            if (FUSED && !arrived && i < F.n_if) {
                const int par = (int)(hseq & 1ULL);
                for (int j = F.if_ptr[i]; j < F.if_ptr[i + 1]; ++j)
                    halo_send_area(F, F.if_nbr[j], par)[3 * (size_t)F.if_pos[j] + r] = yv;
            }
        }
        rb = rbn; re = ren; rbn = rb2; ren = re2; c0 = cn0; c1 = cn1; c2 = cn2;
    }
    if (FUSED && !arrived) arrive();
    const double t = block_sum<THREADS>(dot, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
    if (A.tail) p2p_tail_push<THREADS>(A.tail, &A.tail->tk_spmv, (unsigned)A.tail_total, A.tail_partials, nullptr, A.tail_total, red,
                                       A.tail_out, nullptr, const_cast<CgScalars*>(A.S));
}

// Symmetric-storage product (EN234_SPMV_SYM=1): the reference keeps the upper triangle only (modules/Stiffness.f90:17-28;
// its product visits every stored entry twice, solver_conjugate_gradient.f90:878-883).  Here the assembled block rows stay as
// they are, but a row READS only its diagonal and upper blocks from its own storage (the tail of each of its three scalar
// rows); a lower block (i, j), j < i, is taken as the transpose of block (j, i) from row j's storage — 24 contiguous bytes per
// lane — which row j streamed from HBM a few thousand rows earlier and which is still in L2 when the numbering has a bounded
// front (structured blocks: one node layer).  HBM traffic per row drops from 27 to 14 blocks; every row still sums its own
// terms in a fixed order (gather form, no atomics), so the product stays deterministic.  sym_base[block] = offset of the
// transposed block's first entry (0xffffffff: diagonal / upper block, read in place), sym_len[block] = scalar-row length of
// row j.  Rows of at most 32 blocks only.
template <int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_spmv_sym(SpmvArgs A, const unsigned* __restrict__ sym_base,
                                                            const uint8_t* __restrict__ sym_len) {
    __shared__ double red[WARPS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = gridDim.x * WARPS, N = A.N;
    const int d0 = lane / 3, m0 = lane % 3;
    const int d1 = (lane + 32) / 3, m1 = (lane + 32) % 3;
    const int d2 = (lane + 64) / 3, m2 = (lane + 64) % 3;
    double dot = 0.0;
    int i = A.row0 + blockIdx.x * WARPS + warp;
    int rb = 0, re = 0, rbn = 0, ren = 0, c0 = 0, c1 = 0, c2 = 0;
    unsigned t0 = 0xffffffffu, t1 = 0xffffffffu, t2 = 0xffffffffu;
    int l0 = 0, l1 = 0, l2 = 0;
    if (i < N) { rb = __ldg(A.browptr + i); re = __ldg(A.browptr + i + 1); }
    if (i + W < N) { rbn = __ldg(A.browptr + i + W); ren = __ldg(A.browptr + i + W + 1); }
    {
        const int nb = re - rb;
        if (d0 < nb) { c0 = __ldg(A.bcol + rb + d0); t0 = __ldg(sym_base + rb + d0); l0 = __ldg(sym_len + rb + d0); }
        if (d1 < nb) { c1 = __ldg(A.bcol + rb + d1); t1 = __ldg(sym_base + rb + d1); l1 = __ldg(sym_len + rb + d1); }
        if (d2 < nb) { c2 = __ldg(A.bcol + rb + d2); t2 = __ldg(sym_base + rb + d2); l2 = __ldg(sym_len + rb + d2); }
    }
    pdl_wait();
    pdl_trigger();
    if (A.check_stop && A.S->stop) return;
    for (; i < N; i += W) {
        const int nb = re - rb, L = 3 * nb;
        const double* v = A.val + 9 * (size_t)rb;
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, a00 = 0.0, a01 = 0.0, a02 = 0.0, a10 = 0.0, a11 = 0.0, a12 = 0.0, a20 = 0.0,
               a21 = 0.0, a22 = 0.0;
        if (d0 < nb) {
            x0 = __ldg(A.x + 3 * c0 + m0);
            if (t0 == 0xffffffffu) { a00 = __ldcg(v + lane); a10 = __ldcg(v + L + lane); a20 = __ldcg(v + 2 * L + lane); }
            else { const double* w = A.val + (size_t)t0 + m0 * l0; a00 = __ldcg(w); a10 = __ldcg(w + 1); a20 = __ldcg(w + 2); }
        }
        if (d1 < nb) {
            x1 = __ldg(A.x + 3 * c1 + m1);
            if (t1 == 0xffffffffu) { a01 = __ldcg(v + lane + 32); a11 = __ldcg(v + L + lane + 32); a21 = __ldcg(v + 2 * L + lane + 32); }
            else { const double* w = A.val + (size_t)t1 + m1 * l1; a01 = __ldcg(w); a11 = __ldcg(w + 1); a21 = __ldcg(w + 2); }
        }
        if (d2 < nb) {
            x2 = __ldg(A.x + 3 * c2 + m2);
            if (t2 == 0xffffffffu) { a02 = __ldcg(v + lane + 64); a12 = __ldcg(v + L + lane + 64); a22 = __ldcg(v + 2 * L + lane + 64); }
            else { const double* w = A.val + (size_t)t2 + m2 * l2; a02 = __ldcg(w); a12 = __ldcg(w + 1); a22 = __ldcg(w + 2); }
        }
        const int nbn = ren - rbn;
        int cn0 = 0, cn1 = 0, cn2 = 0, rb2 = 0, re2 = 0, ln0 = 0, ln1 = 0, ln2 = 0;
        unsigned tn0 = 0xffffffffu, tn1 = 0xffffffffu, tn2 = 0xffffffffu;
        if (d0 < nbn) { cn0 = __ldg(A.bcol + rbn + d0); tn0 = __ldg(sym_base + rbn + d0); ln0 = __ldg(sym_len + rbn + d0); }
        if (d1 < nbn) { cn1 = __ldg(A.bcol + rbn + d1); tn1 = __ldg(sym_base + rbn + d1); ln1 = __ldg(sym_len + rbn + d1); }
        if (d2 < nbn) { cn2 = __ldg(A.bcol + rbn + d2); tn2 = __ldg(sym_base + rbn + d2); ln2 = __ldg(sym_len + rbn + d2); }
        if (i + 2 * W < N) { rb2 = __ldg(A.browptr + i + 2 * W); re2 = __ldg(A.browptr + i + 2 * W + 1); }
        const double y0 = a00 * x0 + a01 * x1 + a02 * x2;
        const double y1 = a10 * x0 + a11 * x1 + a12 * x2;
        const double y2 = a20 * x0 + a21 * x1 + a22 * x2;
        const double yv = warp_sum3(y0, y1, y2, lane);
        if ((lane & 7) == 0 && lane < 24) {
            const int r = lane >> 3;
            A.y[3 * (size_t)i + r] = yv;
            dot += yv * A.x[3 * (size_t)i + r];
        }
        rb = rbn; re = ren; rbn = rb2; ren = re2; c0 = cn0; c1 = cn1; c2 = cn2;
        t0 = tn0; t1 = tn1; t2 = tn2; l0 = ln0; l1 = ln1; l2 = ln2;
    }
    const double t = block_sum<THREADS>(dot, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
    if (A.tail) p2p_tail_push<THREADS>(A.tail, &A.tail->tk_spmv, (unsigned)A.tail_total, A.tail_partials, nullptr, A.tail_total, red,
                                       A.tail_out, nullptr, const_cast<CgScalars*>(A.S));
}

struct CgVecArgs {
    int n;                         // number of DOFs
    const double* rhs;
    const double* minv;
    const uint8_t* owned;          // per NODE ownership mask for dot products (null = all owned)
    double *sol, *r, *p, *Ap;
    const double* part_pAp; int n_part_pAp;     // partials written by the operator kernel
    double* part_a; double* part_b; int n_part; // partials written by this family (grid size of the vector kernels)
    CgScalars* S;
    int parity;                    // iteration parity (selects S->rz slot)
    FusedComm F;                   // collectives folded into the vector kernels (F.active; en234gpu_p2p.cuh)
    int tail_finish = 0;           // EN234_P2P_TAIL=2: that CTA also waits for the peers and leaves the totals in part_a[0] / part_b[0]
    ArenaHdr* tail = nullptr;      // separate-kernel collectives: k_cg_update's last CTA pushes the rank's r.r / r.z
};

__device__ __forceinline__ bool dof_owned(const uint8_t* owned, int d) { return owned == nullptr || owned[d / 3] != 0; }

// sol = 0, r = b, p = z = M^-1 r ; partials of b.b and r.z      (:802-847, iteration 1's p = z at :869-870)
__global__ void __launch_bounds__(VEC_THREADS) k_cg_init(CgVecArgs A) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double bb = 0.0, rz = 0.0;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < A.n; d += gridDim.x * VEC_THREADS) {
        const double b = A.rhs[d];
        const double z = b * A.minv[d];
        A.sol[d] = 0.0;
        A.r[d] = b;
        A.p[d] = z;
        if (dof_owned(A.owned, d)) { bb += b * b; rz += z * b; }
    }
    const double s0 = block_sum<VEC_THREADS>(bb, red);
    const double s1 = block_sum<VEC_THREADS>(rz, red);
    if (threadIdx.x == 0) { A.part_a[blockIdx.x] = s0; A.part_b[blockIdx.x] = s1; }
}
// single CTA: bnrm, rz[1], zero-rhs exit (:805-814)
__global__ void __launch_bounds__(VEC_THREADS) k_cg_init_finish(CgVecArgs A) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    const double bb = reduce_partials<VEC_THREADS>(A.part_a, A.n_part, red);
    const double rz = reduce_partials<VEC_THREADS>(A.part_b, A.n_part, red);
    if (threadIdx.x == 0) {
        CgScalars* S = A.S;
        S->bnrm = bb;
        S->rz[1] = rz;
        S->rz[0] = 0.0;
        S->err = 1.0;
        S->iter = 0;
        S->fail = 0;
        S->stop = (bb == 0.0) ? 1 : 0;
    }
}

// ak = (z.r)/(p.Ap); sol += ak p; r -= ak Ap; partials of r.r and r.(M^-1 r)            (:885-900)
// Each thread handles VEC_ILP strided entries with all loads issued before the first use (the vectors are far smaller
// than the matrix, so these kernels are latency-bound unless several loads per thread are in flight).
constexpr int VEC_ILP = 4;
template <bool FUSED>
__global__ void __launch_bounds__(VEC_THREADS, 4) k_cg_update(CgVecArgs A) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    __shared__ double sh2[2];
    const FusedComm& F = A.F;
    const int T = gridDim.x * VEC_THREADS;
    const int dfirst = blockIdx.x * VEC_THREADS + threadIdx.x;
    double pv[VEC_ILP], sv[VEC_ILP], rv[VEC_ILP], av[VEC_ILP], mv[VEC_ILP];
    if (!FUSED) {
        // p, sol, r and the preconditioner were last written at least two kernels ago: with a programmatic launch these
        // loads overlap the tail of the operator kernel; A p and its partial sums are read after pdl_wait()
#pragma unroll
        for (int k = 0; k < VEC_ILP; ++k) {
            const int d = dfirst + k * T;
            if (d < A.n) { pv[k] = A.p[d]; sv[k] = A.sol[d]; rv[k] = A.r[d]; mv[k] = A.minv[d]; }
        }
    }
    pdl_wait();
    pdl_trigger();
    if (A.S->stop) return;
    double pAp;
    unsigned long long seq = 0ULL, hseq = 0ULL;
    ArenaHdr* mine = nullptr;
    if (FUSED) {
        // sequence numbers of this iteration's p.Ap all-reduce and halo exchange
        mine = F.P.arena[F.P.me];
        seq = ld_volatile_u64(&mine->ar_seq) + 1ULL;
        hseq = F.n_nbr > 0 ? ld_volatile_u64(&mine->halo_seq) + 1ULL : 0ULL;
    }
    // CTA 0 exchanges the rank's p.Ap with the peers (and waits for the neighbours' interface rows) before it has anything
    // else in flight; the others pick the all-reduced value up from their local broadcast slot (p2p_allreduce_cta) with
    // the vector loads of their first chunk already under way
    double unused;
    const int nif3 = FUSED ? 3 * F.n_if : 0, hpar = (int)(hseq & 1ULL);
    const bool need_halo = (int)(blockIdx.x * VEC_THREADS) < nif3;     // this CTA's first chunk reaches into the interface DOFs
    if (FUSED && blockIdx.x == 0)
        p2p_allreduce_cta<VEC_THREADS>(F, seq, hseq, A.S, F.pAp_pushed, need_halo, A.part_pAp, A.n_part_pAp, nullptr, 0, red, sh2, pAp, unused);
#pragma unroll
    for (int k = 0; k < VEC_ILP; ++k) {
        const int d = dfirst + k * T;
        if (d < A.n) {
            if (FUSED) { pv[k] = A.p[d]; sv[k] = A.sol[d]; rv[k] = A.r[d]; mv[k] = A.minv[d]; }
            av[k] = A.Ap[d];
        }
    }
    if (FUSED) {
        if (blockIdx.x != 0)
            p2p_allreduce_cta<VEC_THREADS>(F, seq, hseq, A.S, F.pAp_pushed, need_halo, A.part_pAp, A.n_part_pAp, nullptr, 0, red, sh2, pAp, unused);
    } else {
        pAp = reduce_partials<VEC_THREADS>(A.part_pAp, A.n_part_pAp, red);
    }
    const double ak = A.S->rz[A.parity] / pAp;
    double rr = 0.0, rz = 0.0;
    for (int d0 = dfirst; d0 < A.n; d0 += VEC_ILP * T) {
        if (d0 != dfirst) {
#pragma unroll
            for (int k = 0; k < VEC_ILP; ++k) {
                const int d = d0 + k * T;
                if (d < A.n) { pv[k] = A.p[d]; sv[k] = A.sol[d]; rv[k] = A.r[d]; av[k] = A.Ap[d]; mv[k] = A.minv[d]; }
            }
        }
#pragma unroll
        for (int k = 0; k < VEC_ILP; ++k) {
            const int d = d0 + k * T;
            if (d < A.n) {
                double ap = av[k];
                if (FUSED && d < nif3) ap = halo_total(F, hpar, d, ap);          // interface row: add the neighbours' partial sums
                A.sol[d] = sv[k] + ak * pv[k];
                const double r = rv[k] - ak * ap;
                A.r[d] = r;
                if (dof_owned(A.owned, d)) { rr += r * r; rz += (r * mv[k]) * r; }
            }
        }
    }
    const double s0 = block_sum<VEC_THREADS>(rr, red);
    const double s1 = block_sum<VEC_THREADS>(rz, red);
    if (threadIdx.x == 0) { A.part_a[blockIdx.x] = s0; A.part_b[blockIdx.x] = s1; }
    if (!FUSED && A.tail)
        p2p_tail_push<VEC_THREADS>(A.tail, &A.tail->tk_upd, gridDim.x, A.part_a, A.part_b, (int)gridDim.x, red,
                                   A.tail_finish ? A.part_a : nullptr, A.tail_finish ? A.part_b : nullptr, A.S);
    if (FUSED) {
        // last CTA: this all-reduce and halo exchange have been consumed by every CTA -> advance the counters, then reduce
        // the r.r / r.z partials and push them to all ranks (they cross NVLink while k_cg_pupdate is being launched)
        __shared__ int s_last;
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(&mine->tk_upd, 1u) == gridDim.x - 1 ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const double a = reduce_partials_cg<VEC_THREADS>(A.part_a, (int)gridDim.x, red);
            const double b = reduce_partials_cg<VEC_THREADS>(A.part_b, (int)gridDim.x, red);
            if (threadIdx.x == 0) {
                mine->tk_upd = 0;
                mine->ar_seq = seq;
                if (hseq != 0ULL) mine->halo_seq = hseq;
            }
            p2p_push2(F.P, seq + 1ULL, a, b);
        }
    }
}

// err = r.r/b.b; stop if !(err > tol) (:849, :900) or after maxit iterations (:854-860);
// else bk = (z.r)_new/(z.r)_old, p = z + bk p (:867-874)
__global__ void __launch_bounds__(VEC_THREADS) k_cg_pupdate(CgVecArgs A) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    CgScalars* S = A.S;
    const int T = gridDim.x * VEC_THREADS;
    const int dfirst = blockIdx.x * VEC_THREADS + threadIdx.x;
    double pv[VEC_ILP], rv[VEC_ILP], mv[VEC_ILP];
    // p (this kernel's own output of the previous iteration) and the preconditioner are not touched by the predecessor:
    // with a programmatic launch these loads overlap its tail; r and the partial sums are read after pdl_wait()
#pragma unroll
    for (int k = 0; k < VEC_ILP; ++k) {
        const int d = dfirst + k * T;
        if (d < A.n) { pv[k] = A.p[d]; mv[k] = A.minv[d]; }
    }
    pdl_wait();
    pdl_trigger();
    if (S->stop) return;
    double rr, rz;
    unsigned long long seq = 0ULL;
    if (A.F.active) {
        seq = ld_volatile_u64(&A.F.P.arena[A.F.P.me]->ar_seq) + 1ULL;
    }
    // CTA 0 of a partitioned run exchanges r.r / r.z with the peers first; everybody else starts the vector loads of the
    // first chunk before the scalars arrive so their latency overlaps the wait
    __shared__ double sh2[2];
    if (A.F.active && blockIdx.x == 0)
        p2p_allreduce_cta<VEC_THREADS>(A.F, seq, 0ULL, S, 1, false, A.part_a, A.n_part, A.part_b, A.n_part, red, sh2, rr, rz);
#pragma unroll
    for (int k = 0; k < VEC_ILP; ++k) {
        const int d = dfirst + k * T;
        if (d < A.n) rv[k] = A.r[d];
    }
    if (A.F.active) {
        if (blockIdx.x != 0)
            p2p_allreduce_cta<VEC_THREADS>(A.F, seq, 0ULL, S, 1, false, A.part_a, A.n_part, A.part_b, A.n_part, red, sh2, rr, rz);
    } else {
        rr = reduce_partials<VEC_THREADS>(A.part_a, A.n_part, red);
        rz = reduce_partials<VEC_THREADS>(A.part_b, A.n_part, red);
    }
    const double err = rr / S->bnrm;
    const bool more = err > S->tol;
    const int iter = S->iter + 1;
    const bool exhausted = more && iter >= S->maxit;
    const double bk = rz / S->rz[A.parity];
    if (more && !exhausted) {
#pragma unroll
        for (int k = 0; k < VEC_ILP; ++k) {
            const int d = dfirst + k * T;
            if (d < A.n) A.p[d] = bk * pv[k] + rv[k] * mv[k];
        }
        for (int d0 = dfirst + VEC_ILP * T; d0 < A.n; d0 += VEC_ILP * T) {
#pragma unroll
            for (int k = 0; k < VEC_ILP; ++k) {
                const int d = d0 + k * T;
                if (d < A.n) { pv[k] = A.p[d]; rv[k] = A.r[d]; mv[k] = A.minv[d]; }
            }
#pragma unroll
            for (int k = 0; k < VEC_ILP; ++k) {
                const int d = d0 + k * T;
                if (d < A.n) A.p[d] = bk * pv[k] + rv[k] * mv[k];
            }
        }
    }
    // bookkeeping by the LAST CTA to finish reading S (all CTAs read S->iter/stop above): use a grid-wide ticket
    __shared__ unsigned ticket;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ticket = atomicAdd((unsigned*)&S->pad0, 1u);
    }
    __syncthreads();
    if (ticket == gridDim.x - 1 && threadIdx.x == 0) {
        S->pad0 = 0;
        if (A.F.active) A.F.P.arena[A.F.P.me]->ar_seq = seq;
        S->rz[A.parity ^ 1] = rz;
        S->err = err;
        S->iter = iter;
        if (!more) S->stop = 1;
        if (exhausted) { S->stop = 1; S->fail = 1; }
        __threadfence();
    }
}

// dof_increment += sol ; partial of sol.sol        (solve_conjugategradient :760-771)
__global__ void __launch_bounds__(VEC_THREADS) k_add_solution(int n, const double* sol, double* du, const uint8_t* owned,
                                                               double* partials) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double ss = 0.0;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        const double s = sol[d];
        du[d] = du[d] + s;
        if (dof_owned(owned, d)) ss += s * s;
    }
    const double t = block_sum<VEC_THREADS>(ss, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

}  // namespace en234k
