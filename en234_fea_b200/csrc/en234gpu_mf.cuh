// Matrix-free element-by-element operator  y = sum_e P_e^T K_e P_e x  (BASELINE config 5; the reference has no such
// path — its CG SpMV, src/solver_conjugate_gradient.f90:877-883, always reads an assembled matrix).
// K_e x_e is evaluated as the internal force of the linear-elastic element for the displacement field x
// (the integrand of el_linelast_3Dbasic.f90:124-128 with u := x), never forming K_e.
// Operator application (mf_apply): the residual-only element kernel of the assembly (k_element_forces, reading the
// masked vector x instead of the DOFs) leaves f_e = -K_e x_e in the node-major element scratch and
// k_mf_gather sums every node's contributions in ascending element order — no atomics, no colouring, deterministic.
// The Jacobi diagonal (mf_diag, once per boundary-condition pass) is still integrated colour by colour: elements are
// greedily coloured so that no two elements of a colour share a node, colours are processed by successive launches.
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "en234gpu_cg.cuh"
#include "en234gpu_elem.cuh"

namespace en234k {

// the dominant distinct row (27 blocks: an interior node of a hex mesh), passed BY VALUE: kernel parameters live in the
// constant bank, so the fully unrolled fast path uses its 243 values as immediate constant operands of the FMAs —
// no value loads at all
struct DominantRow { double v[243]; };

struct MfData {
    double* d_fe = nullptr;         // element force scratch, node-major (3 per adjacency entry)
    // Gauss-point geometry cache of k_element_forces (J^-1 and w|J|: 640 B per element).  The mesh never moves, so it is
    // written by the first launch of the kernel and read by all later ones.  EN234_MF_GEOM=0 disables it (A/B).
    double* d_geom = nullptr;
    int geom_state = 0;             // 0 not allocated yet, 1 allocated and to be written, 2 valid, -1 disabled / no memory
    int grid_cached = 0, occ_cached = 2;   // persistent grid / CTAs per SM of k_element_forces_cached (EN234_MF_OCC=3: A/B)
    int sm_share = 1;               // > 1: several parts of one mesh share the device (their kernels must be co-resident)
    bool tpe = true;                // thread-per-element kernel on the transposed cache (EN234_MF_TPE=0: warp-cooperative forms)
    size_t geom_stride = 0;
    // congruent-mesh fast path: every element is a translate of element 0 (bitwise equal node offsets, same material),
    // so one 24x24 K_e serves the whole mesh and a block row is determined by WHICH local nodes of its adjacent elements
    // the node is (and where their columns sit in the row): a few dozen distinct rows on a structured block.  The
    // operator is then a block-CSR SpMV whose values come from the table of distinct rows (k_stencil_nodes): no global
    // matrix, HBM traffic = column indices + vectors.
    bool congruent = false;
    double* d_st_val = nullptr;     // concatenated distinct rows, [r][block][c] each
    int* d_st_off = nullptr;        // start of each distinct row in d_st_val
    int* d_st_pid = nullptr;        // distinct-row id of every block row
    int* d_st_nb = nullptr;         // blocks per distinct row
    int* d_perm = nullptr;          // pattern-sorted row order
    int n_dom = 0;                  // rows of the dominant distinct row (first in the sorted order)
    DominantRow dom;
    int* d_bcolT = nullptr;         // ELL-transposed block columns in sorted order (coalesced for one thread per row)
    double* d_xs = nullptr;         // masked, component-major copy of the operand
    int n_patterns = 0;
    int t_split = 0, grid_tail = 0, grid_head = 0;
    const uint8_t* affected = nullptr;   // per node: prescribed or coupled to a prescribed DOF (set with the BC list)
    int grid_elem = 0, elem_threads = 0, grid_gather = 0;
    size_t elem_smem = 0;
    int n_colors = 0;
    std::vector<int> color_ptr;     // host
    int* d_elist = nullptr;         // elements grouped by colour, ascending within a colour
    int n_partials = 0;
    int launches_per_apply = 0;
    int grid_vec = 0;
};

// 8 lanes = the 8 Gauss points of one element.  mode 1: diag(K_e) into y (the Jacobi diagonal of the matrix-free solve,
// once per boundary-condition pass);  mode 0: y += K_e x — the colour-pass operator the element-kernel + gather form
// replaced, kept for A/B.
template <int MODE>
__global__ void __launch_bounds__(256) k_mf_color(Mesh M, int n, const int* __restrict__ elist, const uint8_t* cflag,
                                                  const double* __restrict__ x, double* __restrict__ y,
                                                  const CgScalars* S) {
    if (S && S->stop) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int idx = t >> 3, gp = t & 7;
    const bool valid = idx < n;
    const int e = elist[valid ? idx : 0];
    double g[GEOM_STRIDE];
    GpOut o;
    gauss_point_eval(M, e, gp, LoadMasked{x, cflag}, g, o);
    const Material mat = M.mats[M.mat[e]];
    double f[24];
    if (MODE == 0) {
        double s[9];
        stress_linear(mat, o.H, s);
#pragma unroll
        for (int q = 0; q < 9; ++q) s[q] *= o.det;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            f[3 * a] = s[0] * g[3 * a] + s[1] * g[3 * a + 1] + s[2] * g[3 * a + 2];
            f[3 * a + 1] = s[3] * g[3 * a] + s[4] * g[3 * a + 1] + s[5] * g[3 * a + 2];
            f[3 * a + 2] = s[6] * g[3 * a] + s[7] * g[3 * a + 1] + s[8] * g[3 * a + 2];
        }
    } else {
        const double lm = (mat.d12 + mat.d44) * o.det, mu = mat.d44 * o.det;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const double n2 = mu * (g[3 * a] * g[3 * a] + g[3 * a + 1] * g[3 * a + 1] + g[3 * a + 2] * g[3 * a + 2]);
#pragma unroll
            for (int i = 0; i < 3; ++i) f[3 * a + i] = lm * g[3 * a + i] * g[3 * a + i] + n2;
        }
    }
    // halving butterfly over the 8 Gauss-point lanes: lane a ends with the 3 components of node a
    double h[12];
    {
        const bool up = gp & 4;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const double send = up ? f[j] : f[12 + j];
            const double keep = up ? f[12 + j] : f[j];
            h[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    double q6[6];
    {
        const bool up = gp & 2;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double send = up ? h[j] : h[6 + j];
            const double keep = up ? h[6 + j] : h[j];
            q6[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    double r3[3];
    {
        const bool up = gp & 1;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double send = up ? q6[j] : q6[3 + j];
            const double keep = up ? q6[3 + j] : q6[j];
            r3[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
    if (valid) {
        const int node = M.conn[(size_t)gp * M.E + e];     // lane a owns local node a
        double* yy = y + 3 * (size_t)node;
        yy[0] += r3[0]; yy[1] += r3[1]; yy[2] += r3[2];
    }
}

__global__ void k_mask_values(int n, const uint8_t* cflag, const double* cval, double* out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) out[d] = cflag[d] ? cval[d] : 0.0;
}
__global__ void k_fix_diag(int n, const uint8_t* cflag, const uint8_t* owned, double* diag) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n && cflag[d]) diag[d] = (owned == nullptr || owned[d / 3]) ? 1.0 : 0.0;
}

struct StencilArgs {
    int N;
    int t_begin, t_end;            // sorted positions handled by this launch
    int n_dom;                     // the first n_dom rows of the pattern-sorted order all have the dominant distinct row
    const int* perm;               // pattern-sorted position t -> block row
    const int* bcolT;              // ELL-transposed block columns in sorted order: bcolT[s * N + t] = s-th column of row perm[t]
    const double* st_val;          // concatenated distinct rows, [r][block][c] each
    const int* st_off;             // start of each distinct row
    const int* st_nb;              // number of blocks of each distinct row
    const int* st_pid;             // distinct-row id per sorted position
    const uint8_t* cflag;          // per DOF: 1 = prescribed (null: none)
    const uint8_t* affected;       // per node: row is prescribed or couples to a prescribed DOF (null: test cflag always)
    const uint8_t* owned;
    const double* x;
    const double* xs;              // masked x, component-major: xs[c * N + node] (written by k_mask_soa)
    double* y;
    double* partials;
    const CgScalars* S;
};

// xs[c][node] = P x: prescribed DOFs zeroed, AoS (u1 u2 u3 per node, the reference layout) -> component-major, so that
// the per-thread gathers of k_stencil_nodes touch 8-byte-stride sectors instead of 24-byte-stride ones
__global__ void k_mask_soa(int N, const uint8_t* cflag, const double* __restrict__ x, double* __restrict__ xs,
                           const CgScalars* S) {
    if (S && S->stop) return;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= 3 * N) return;
    xs[(size_t)(d % 3) * N + d / 3] = (cflag && cflag[d]) ? 0.0 : x[d];
}

// y = (P A P + (I - P)) x for a congruent mesh, ONE THREAD PER NODE ROW, rows visited in pattern-sorted order so that a
// warp's 32 rows share one distinct row.  (The warp-per-row SpMV spends ~170 warp instructions per row on index
// pipelining and three warp reductions — fine while 2 kB of matrix values stream past per row, but that is all that is
// left once the values come from a table.)  A thread walks its row's blocks: column index from the ELL-transposed index
// array (coalesced across the warp), three x values, nine FMAs.  Dominant-row CTAs take the unrolled constant-operand
// path; the others read their values from the table (same address for every lane of a warp: one L1 sector per load).
#ifndef STENCIL_CHUNK
#define STENCIL_CHUNK 9
#endif
#ifndef STENCIL_OCC
#define STENCIL_OCC 3
#endif
__global__ void __launch_bounds__(VEC_THREADS, STENCIL_OCC) k_stencil_nodes(StencilArgs A, const __grid_constant__ DominantRow D) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (A.S && A.S->stop) return;
    double dot = 0.0;
    const int N = A.N;
    for (int t0 = A.t_begin + blockIdx.x * VEC_THREADS; t0 < A.t_end; t0 += gridDim.x * VEC_THREADS) {
        const int t = t0 + threadIdx.x;
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
        if (t0 + VEC_THREADS <= A.n_dom) {                       // CTA-uniform: 256 dominant rows
            const int* bc = A.bcolT + t;
            const double* xs0 = A.xs;
            const double* xs1 = A.xs + (size_t)N;
            const double* xs2 = A.xs + 2 * (size_t)N;
            int col[27];                                         // all column indices first: one HBM latency per row
#pragma unroll
            for (int s = 0; s < 27; ++s) col[s] = __ldg(bc + (size_t)s * N);
#pragma unroll
            for (int s0 = 0; s0 < 27; s0 += STENCIL_CHUNK) {
                double xv[STENCIL_CHUNK][3];
#pragma unroll
                for (int u = 0; u < STENCIL_CHUNK; ++u) {
                    xv[u][0] = __ldg(xs0 + col[s0 + u]); xv[u][1] = __ldg(xs1 + col[s0 + u]); xv[u][2] = __ldg(xs2 + col[s0 + u]);
                }
#pragma unroll
                for (int u = 0; u < STENCIL_CHUNK; ++u) {
                    const int s = s0 + u;
                    const double x0 = xv[u][0], x1 = xv[u][1], x2 = xv[u][2];
                    y0 += D.v[3 * s] * x0;       y0 += D.v[3 * s + 1] * x1;       y0 += D.v[3 * s + 2] * x2;
                    y1 += D.v[81 + 3 * s] * x0;  y1 += D.v[81 + 3 * s + 1] * x1;  y1 += D.v[81 + 3 * s + 2] * x2;
                    y2 += D.v[162 + 3 * s] * x0; y2 += D.v[162 + 3 * s + 1] * x1; y2 += D.v[162 + 3 * s + 2] * x2;
                }
            }
        } else if (t < A.t_end) {
            const int pid = A.st_pid[t], nb = A.st_nb[pid], L = 3 * nb;
            const double* v = A.st_val + A.st_off[pid];
            for (int s0 = 0; s0 < nb; s0 += 9) {
                int col[9];
#pragma unroll
                for (int u = 0; u < 9; ++u) col[u] = s0 + u < nb ? __ldg(A.bcolT + (size_t)(s0 + u) * N + t) : 0;
                double xv[9][3];
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    xv[u][0] = __ldg(A.xs + col[u]); xv[u][1] = __ldg(A.xs + (size_t)N + col[u]); xv[u][2] = __ldg(A.xs + 2 * (size_t)N + col[u]);
                }
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    if (s0 + u < nb) {
                        const double* vs = v + 3 * (s0 + u);
                        y0 += __ldg(vs) * xv[u][0];         y0 += __ldg(vs + 1) * xv[u][1];         y0 += __ldg(vs + 2) * xv[u][2];
                        y1 += __ldg(vs + L) * xv[u][0];     y1 += __ldg(vs + L + 1) * xv[u][1];     y1 += __ldg(vs + L + 2) * xv[u][2];
                        y2 += __ldg(vs + 2 * L) * xv[u][0]; y2 += __ldg(vs + 2 * L + 1) * xv[u][1]; y2 += __ldg(vs + 2 * L + 2) * xv[u][2];
                    }
                }
            }
        }
        if (t < A.t_end) {
            const int i = A.perm[t];
            const bool aff = A.cflag != nullptr && (A.affected == nullptr || A.affected[i] != 0);
            double yv[3] = {y0, y1, y2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t d = 3 * (size_t)i + c;
                const double xv = A.x[d];
                if (aff && A.cflag[d]) yv[c] = (A.owned == nullptr || A.owned[i]) ? xv : 0.0;
                A.y[d] = yv[c];
                dot += yv[c] * xv;
            }
        }
    }
    const double tsum = block_sum<VEC_THREADS>(dot, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = tsum;
}

// y_i = -sum over the node's (element, local node) entries of the element scratch, ascending element order;
// y_row = x_row on prescribed rows (unit diagonal on the owner rank only); partial of x.y over all local rows.
// Eight lanes per node: lane k fetches adjacency entry k (all fetches of a node in flight together), the three
// components are then summed across the eight lanes in ascending k by a shuffle chain (fixed order: deterministic).
__global__ void __launch_bounds__(VEC_THREADS) k_mf_gather(int N, const int* __restrict__ adj_ptr, const int* __restrict__ adj,
                                                            const double* __restrict__ fe, const uint8_t* cflag,
                                                            const uint8_t* owned, const double* __restrict__ x, double* y,
                                                            double* partials, const CgScalars* S) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S && S->stop) return;
    double dot = 0.0;
    const int lane = threadIdx.x & 31, k = lane & 7, base = lane & 24;
    const int per_cta = VEC_THREADS / 8;
    for (int i0 = blockIdx.x * per_cta; i0 < N; i0 += gridDim.x * per_cta) {
        const int i = i0 + (threadIdx.x >> 3);
        const bool live = i < N;
        const int k0 = live ? adj_ptr[i] : 0, k1 = live ? adj_ptr[i + 1] : 0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        int kmax = k1 - k0;                                       // warp-uniform trip count: the longest of the 4 rows
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, 8));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, 16));
        for (int kk = 0; kk < kmax; kk += 8) {
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            if (k0 + kk + k < k1) {
                const double* f = fe + 3 * (size_t)(k0 + kk + k);     // node-major scratch: contiguous per node
                v0 = __ldcs(f); v1 = __ldcs(f + 1); v2 = __ldcs(f + 2);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {                         // padding entries add 0
                s0 += __shfl_sync(0xffffffffu, v0, base + u);
                s1 += __shfl_sync(0xffffffffu, v1, base + u);
                s2 += __shfl_sync(0xffffffffu, v2, base + u);
            }
        }
        if (live && k < 3) {
            const size_t d = 3 * (size_t)i + k;
            double v = -(k == 0 ? s0 : (k == 1 ? s1 : s2));
            const double xv = x[d];
            if (cflag && cflag[d]) v = (owned == nullptr || owned[i]) ? xv : 0.0;
            y[d] = v;
            dot += v * xv;
        }
    }
    const double t = block_sum<VEC_THREADS>(dot, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

inline int mf_setup(MfData& mf, const Mesh& M, const std::vector<int>& conn /*SoA, host*/, int N, int E, int n_sm, int ndof) {
    // greedy colouring in element order: colour = smallest not used by an earlier element sharing a node
    std::vector<unsigned long long> node_mask(N, 0ULL);
    std::vector<int> color(E);
    int ncol = 0;
    for (int e = 0; e < E; ++e) {
        unsigned long long used = 0ULL;
        for (int a = 0; a < 8; ++a) used |= node_mask[conn[(size_t)a * E + e]];
        int c = 0;
        while (c < 64 && (used >> c) & 1ULL) ++c;
        if (c >= 64) return 1;
        color[e] = c;
        ncol = std::max(ncol, c + 1);
        for (int a = 0; a < 8; ++a) node_mask[conn[(size_t)a * E + e]] |= 1ULL << c;
    }
    mf.n_colors = ncol;
    mf.color_ptr.assign(ncol + 1, 0);
    for (int e = 0; e < E; ++e) mf.color_ptr[color[e] + 1]++;
    for (int c = 0; c < ncol; ++c) mf.color_ptr[c + 1] += mf.color_ptr[c];
    std::vector<int> elist(E), pos(mf.color_ptr.begin(), mf.color_ptr.end() - 1);
    for (int e = 0; e < E; ++e) elist[pos[color[e]]++] = e;
    if (cudaMalloc((void**)&mf.d_elist, (size_t)E * sizeof(int)) != cudaSuccess) return 1;
    cudaMemcpy(mf.d_elist, elist.data(), (size_t)E * sizeof(int), cudaMemcpyHostToDevice);
    mf.grid_vec = std::min(MAX_PARTIALS, std::max(1, std::min((ndof + VEC_THREADS - 1) / VEC_THREADS, n_sm * 8)));
    // operator application: element kernel (one warp per four elements, persistent) + node gather
    if (cudaMalloc((void**)&mf.d_fe, (size_t)E * 24 * sizeof(double)) != cudaSuccess) return 1;
    mf.elem_smem = 0;
    mf.elem_threads = 256;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_element_forces<LoadMasked>, mf.elem_threads, 0) != cudaSuccess) return 1;
    const int wpc = mf.elem_threads / 32, groups = (E + 3) / 4;
    mf.grid_elem = std::max(1, std::min((groups + wpc - 1) / wpc, n_sm * std::max(occ, 1)));
    mf.grid_gather = std::min(MAX_PARTIALS, std::max(1, (N + VEC_THREADS / 8 - 1) / (VEC_THREADS / 8)));
    mf.n_partials = mf.grid_gather;
    mf.launches_per_apply = 2;
    (void)M;
    return 0;
}
// Table of distinct block rows of a congruent mesh from the K_e of the representative element (reference layout
// element_stiffness(ROW,COL), column-major 24x24).  Each row is summed over its adjacent elements in ascending element
// order — exactly what the assembly produces when every element has this K_e.  Returns 1 on failure, 2 when the mesh
// has too many distinct rows for the table to pay off (the caller then keeps the element-by-element operator).
inline int mf_set_congruent(MfData& mf, const double* Ke_colmajor, int N, const std::vector<int>& adj_ptr,
                            const std::vector<int>& adj, const std::vector<uint8_t>& slot, const std::vector<int>& browptr,
                            const std::vector<int>& bcol) {
    std::map<std::string, int> ids;
    std::vector<int> pid(N), off;
    std::vector<double> val;
    std::string key;
    for (int i = 0; i < N; ++i) {
        const int k0 = adj_ptr[i], k1 = adj_ptr[i + 1], nb = browptr[i + 1] - browptr[i], L = 3 * nb;
        if (nb > 32) return 2;                                   // k_spmv_stencil keeps at most 96 scalar columns per row
        key.assign(1, (char)nb);
        for (int k = k0; k < k1; ++k) {
            key.push_back((char)(adj[k] & 7));
            key.append(reinterpret_cast<const char*>(&slot[(size_t)k * 8]), 8);
        }
        auto it = ids.find(key);
        if (it == ids.end()) {
            if (ids.size() >= 4096) return 2;
            it = ids.emplace(key, (int)ids.size()).first;
            off.push_back((int)val.size());
            val.resize(val.size() + (size_t)9 * nb, 0.0);
            double* row = val.data() + off.back();
            for (int k = k0; k < k1; ++k) {
                const int a = adj[k] & 7;
                for (int b = 0; b < 8; ++b) {
                    const int sl = slot[(size_t)k * 8 + b];
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) row[r * L + 3 * sl + c] += Ke_colmajor[(3 * a + r) + 24 * (3 * b + c)];
                }
            }
        }
        pid[i] = it->second;
    }
    mf.n_patterns = (int)ids.size();
    // pattern-sorted row order: the most frequent distinct row with 27 blocks first (unrolled fast path), then by id
    std::vector<long long> freq(ids.size(), 0);
    for (int i = 0; i < N; ++i) freq[pid[i]]++;
    std::vector<int> nbp(ids.size());
    for (int i = 0; i < N; ++i) nbp[pid[i]] = browptr[i + 1] - browptr[i];
    int dom = -1;
    for (size_t q = 0; q < ids.size(); ++q)
        if (nbp[q] == 27 && (dom < 0 || freq[q] > freq[dom])) dom = (int)q;
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
        const int ka = pid[a] == dom ? -1 : pid[a], kb = pid[b] == dom ? -1 : pid[b];
        return ka < kb;
    });
    mf.n_dom = dom >= 0 ? (int)freq[dom] : 0;
    std::memset(mf.dom.v, 0, sizeof(mf.dom.v));
    if (dom >= 0) std::memcpy(mf.dom.v, val.data() + off[dom], 243 * sizeof(double));
    std::vector<int> pidp(N);
    int width = 0;
    for (int t = 0; t < N; ++t) { pidp[t] = pid[perm[t]]; width = std::max(width, nbp[pidp[t]]); }
    std::vector<int> bt((size_t)width * N, 0);
    for (int t = 0; t < N; ++t) {
        const int i = perm[t];
        for (int sidx = 0; sidx < browptr[i + 1] - browptr[i]; ++sidx) bt[(size_t)sidx * N + t] = bcol[(size_t)browptr[i] + sidx];
    }
    auto up = [](void** d, const void* h, size_t bytes) {
        if (cudaMalloc(d, bytes) != cudaSuccess) return 1;
        return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 1;
    };
    if (up((void**)&mf.d_st_val, val.data(), val.size() * sizeof(double))) return 1;
    if (up((void**)&mf.d_st_off, off.data(), off.size() * sizeof(int))) return 1;
    if (up((void**)&mf.d_st_nb, nbp.data(), nbp.size() * sizeof(int))) return 1;
    if (up((void**)&mf.d_st_pid, pidp.data(), (size_t)N * sizeof(int))) return 1;
    if (up((void**)&mf.d_perm, perm.data(), (size_t)N * sizeof(int))) return 1;
    if (up((void**)&mf.d_bcolT, bt.data(), bt.size() * sizeof(int))) return 1;
    if (cudaMalloc((void**)&mf.d_xs, (size_t)3 * N * sizeof(double)) != cudaSuccess) return 1;
    // two launches per application: the rows after the dominant run first (they include every partition-interface row,
    // whose halo exchange then overlaps the dominant rows), then the dominant run
    mf.t_split = (mf.n_dom / VEC_THREADS) * VEC_THREADS;
    mf.grid_tail = std::min(MAX_PARTIALS / 2, std::max(1, (N - mf.t_split + VEC_THREADS - 1) / VEC_THREADS));
    mf.grid_head = mf.t_split > 0 ? std::min(MAX_PARTIALS / 2, mf.t_split / VEC_THREADS) : 0;
    mf.n_partials = mf.grid_tail + mf.grid_head;
    mf.launches_per_apply = 3;
    mf.congruent = true;
    return 0;
}
// Allocation of the geometry cache (never inside a stream capture: callers that capture call this first).
inline void mf_prepare_geom(MfData& mf, int E) {
    if (mf.geom_state != 0) return;
    const char* e = getenv("EN234_MF_GEOM");
    const char* t = getenv("EN234_MF_TPE");
    mf.tpe = !(t && atoi(t) == 0);
    mf.geom_stride = mf.tpe ? (size_t)((E + 3) / 4) * 4 : 0;
    const size_t bytes = (size_t)((E + 3) / 4) * 320 * sizeof(double);
    if ((e && atoi(e) == 0) || cudaMalloc((void**)&mf.d_geom, bytes) != cudaSuccess) { cudaGetLastError(); mf.geom_state = -1; }
    else mf.geom_state = 1;
}
// Launch of the residual-only element kernel with the geometry cache in the right state.
template <class Load>
inline void launch_element_forces(MfData& mf, ElemArgs EA, const Load& ld, int grid, cudaStream_t s) {
    // the cache pays for the operator of the matrix-free solve (thousands of applications per step, measured 0.89 -> 0.75 ms
    // at 128^3); the residual assembly and the explicit-dynamics step (LoadSum, once per assembly / time step) measured
    // slower with it (64^3: 181 vs 128 us per dynamic step) and keep the coordinate-based kernel
    if (!Load::use_geometry_cache) { k_element_forces<Load, 0><<<grid, 256, 0, s>>>(EA, ld); return; }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (mf.geom_state == 0 && cap == cudaStreamCaptureStatusNone) mf_prepare_geom(mf, EA.M.E);
    EA.geom = mf.d_geom;
    EA.geom_stride = mf.geom_stride;
    if (mf.geom_state == 0) { k_element_forces<Load, 0><<<grid, 256, 0, s>>>(EA, ld); return; }   // capturing, no cache yet
    if (mf.geom_state == 2 && mf.tpe) {
        if (mf.grid_cached == 0) {
            int occ = 1, dev = 0, n_sm = 148;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_element_forces_tpe<Load>, TPE_THREADS, 0);
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            mf.grid_cached = std::max(1, std::min((EA.M.E + TPE_THREADS - 1) / TPE_THREADS, std::max(1, n_sm / mf.sm_share) * std::max(occ, 1)));
        }
        k_element_forces_tpe<Load><<<mf.grid_cached, TPE_THREADS, 0, s>>>(EA, ld);
    } else if (mf.geom_state == 2) {
        if (mf.grid_cached == 0) {
            const char* o = getenv("EN234_MF_OCC");
            mf.occ_cached = (o && atoi(o) == 3) ? 3 : 2;
            int occ = 1;
            if (mf.occ_cached == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_element_forces_cached<Load, 3>, 256, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_element_forces_cached<Load, 2>, 256, 0);
            int dev = 0, n_sm = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            mf.grid_cached = std::max(1, std::min(((EA.M.E + 3) / 4 + 7) / 8, std::max(1, n_sm / mf.sm_share) * std::max(occ, 1)));
        }
        const char* v = getenv("EN234_MF_XI");
        if (v && atoi(v) == 0) k_element_forces<Load, 2><<<grid, 256, 0, s>>>(EA, ld);
        else if (mf.occ_cached == 3) k_element_forces_cached<Load, 3><<<mf.grid_cached, 256, 0, s>>>(EA, ld);
        else k_element_forces_cached<Load, 2><<<mf.grid_cached, 256, 0, s>>>(EA, ld);
    }
    else if (mf.geom_state == 1) {
        k_element_forces<Load, 1><<<grid, 256, 0, s>>>(EA, ld);
        // a kernel whose Load can make it return at once (PCG stop flag) may not have written the cache
        if (cap == cudaStreamCaptureStatusNone && !ld.may_skip()) mf.geom_state = 2;
    } else k_element_forces<Load, 0><<<grid, 256, 0, s>>>(EA, ld);
}

inline void mf_release(MfData& mf) {
    if (mf.d_geom) cudaFree(mf.d_geom);
    mf.d_geom = nullptr; mf.geom_state = 0;
    if (mf.d_elist) cudaFree(mf.d_elist);
    if (mf.d_fe) cudaFree(mf.d_fe);
    if (mf.d_st_val) cudaFree(mf.d_st_val);
    if (mf.d_st_off) cudaFree(mf.d_st_off);
    if (mf.d_st_pid) cudaFree(mf.d_st_pid);
    if (mf.d_bcolT) cudaFree(mf.d_bcolT);
    if (mf.d_xs) cudaFree(mf.d_xs);
    if (mf.d_st_nb) cudaFree(mf.d_st_nb);
    if (mf.d_perm) cudaFree(mf.d_perm);
    mf.d_bcolT = nullptr; mf.d_xs = nullptr; mf.d_st_nb = nullptr; mf.d_perm = nullptr;
    mf.d_elist = nullptr; mf.d_fe = nullptr; mf.d_st_val = nullptr; mf.d_st_off = nullptr; mf.d_st_pid = nullptr;
}

// congruent-mesh operator in two phases: 0 = masked operand copy + the rows after the dominant run, 1 = the dominant run
inline void mf_stencil(const MfData& mf, const Mesh& M, const uint8_t* cflag, const uint8_t* owned, const double* x, double* y,
                       double* partials, const CgScalars* S, cudaStream_t s, long long* launches, int phase) {
    StencilArgs A;
    A.N = M.N; A.n_dom = mf.n_dom; A.perm = mf.d_perm; A.bcolT = mf.d_bcolT; A.st_val = mf.d_st_val;
    A.st_off = mf.d_st_off; A.st_nb = mf.d_st_nb; A.st_pid = mf.d_st_pid; A.cflag = cflag;
    A.affected = cflag ? mf.affected : nullptr; A.owned = owned;
    A.x = x; A.xs = mf.d_xs; A.y = y; A.S = S;
    if (phase == 0) {
        k_mask_soa<<<(3 * M.N + 255) / 256, 256, 0, s>>>(M.N, cflag, x, mf.d_xs, S);
        A.t_begin = mf.t_split; A.t_end = M.N; A.partials = partials;
        k_stencil_nodes<<<mf.grid_tail, VEC_THREADS, 0, s>>>(A, mf.dom);
        *launches += 2;
    } else if (mf.grid_head > 0) {
        A.t_begin = 0; A.t_end = mf.t_split; A.partials = partials + mf.grid_tail;
        k_stencil_nodes<<<mf.grid_head, VEC_THREADS, 0, s>>>(A, mf.dom);
        *launches += 1;
    }
}

// y = P A P x + (I-P) x ; partials of x.y.  (cflag == null: plain A x)
inline void mf_apply(MfData& mf, const Mesh& M, int ndof, const uint8_t* cflag, const uint8_t* owned, const double* x,
                     double* y, double* partials, const CgScalars* S, cudaStream_t s, long long* launches) {
    (void)ndof;
    ElemArgs EA;
    EA.M = M; EA.Ke = nullptr; EA.Re = mf.d_fe; EA.S = nullptr;
    if (mf.congruent) {
        mf_stencil(mf, M, cflag, owned, x, y, partials, S, s, launches, 0);
        mf_stencil(mf, M, cflag, owned, x, y, partials, S, s, launches, 1);
        return;
    }
    launch_element_forces(mf, EA, LoadMasked{x, cflag, S}, mf.grid_elem, s);
    k_mf_gather<<<mf.grid_gather, VEC_THREADS, 0, s>>>(M.N, M.adj_ptr, M.adj, mf.d_fe, cflag, owned, x, y, partials, S);
    *launches += 2;
}
inline void mf_diag(const MfData& mf, const Mesh& M, int ndof, const double* any_vec, double* diag, cudaStream_t s,
                    long long* launches) {
    cudaMemsetAsync(diag, 0, (size_t)ndof * sizeof(double), s);
    for (int c = 0; c < mf.n_colors; ++c) {
        const int n = mf.color_ptr[c + 1] - mf.color_ptr[c];
        if (n == 0) continue;
        k_mf_color<1><<<(8 * n + 255) / 256, 256, 0, s>>>(M, n, mf.d_elist + mf.color_ptr[c], nullptr, any_vec, diag, nullptr);
    }
    *launches += mf.n_colors;
}

}  // namespace en234k
