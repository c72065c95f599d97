// Two-kernel assembly: (1) batched element kernel  (2) deterministic gather into the block-CSR rows.
//
//   k_element_blocks<NEO>   element loop of assemble_*_stiffness (src/solver_direct.f90:102-258): every element is
//                           evaluated ONCE — Gauss-point geometry, stress, the 24-entry element residual and the 64
//                           3x3 blocks of element_stiffness — and written to an element scratch array in HBM.
//   k_gather_rows           the scatter of src/solver_direct.f90:264-287 turned into a gather: one warp per node row
//                           sums the row's (element, local node) contributions in ASCENDING ELEMENT ORDER through the
//                           precomputed element-to-CSR slot map, so every matrix entry and residual entry is summed
//                           in the reference's serial order — no atomics, bitwise run-to-run deterministic.
//
// Element scratch layout (per pair of elements, 2 x 576 doubles): values j = 2jp, 2jp+1 (jp = 0..17) of lane l (0..31)
// at [jp*64 + 2l .. +1], lane l = (e&1)*16 + tA*4 + tB holding the 2x2 tile of blocks {2tA, 2tA+1} x {2tB, 2tB+1},
// j = (sa*2 + sb)*9 + 3r + c.  Kernel 1 therefore stores 512-byte rows with 16-byte stores (fully coalesced) and the
// 72 values of one (element, local node) row are nine whole 64-byte chunks for kernel 2.
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

struct ElemArgs {
    Mesh M;
    double* Ke;             // element blocks, layout above; padded to a multiple of 4 elements
    double* Re;             // element residual entries in NODE-MAJOR order: the three components of (element e, local node
                            // a) at 3 * adj_pos[8e + a], so a node's contributions are contiguous, ascending element order
    CgScalars* S;           // nanflag (may be null)
    double* geom = nullptr; // cached Gauss-point geometry of k_element_forces: per group of 4 elements 10 rows of 32 lanes
                            // (lane = 8 * element-in-group + Gauss point): J^-1 (9 values) and w|J|; 640 B per element
    size_t geom_stride = 0; // != 0: transposed cache layout geomT[(10 * point + q) * geom_stride + e] (k_element_forces_tpe)
};

template <bool NEO>
struct ElemCfg {
    static constexpr int NW = NEO ? 4 : 8;                                      // warps per CTA
    static constexpr int GS = (NEO ? GEOM_STRIDE_NEO : GEOM_STRIDE) + 10;       // + 9 stress values, odd stride
    static constexpr int G_SIG = NEO ? GEOM_STRIDE_NEO : GEOM_STRIDE;           // w * stress (9)
    static constexpr int NSTRIDE = 49;                                          // node buffer: 8 nodes x (xyz, u) + pad
    static constexpr int PER_WARP = 4 * 8 * GS + 4 * NSTRIDE;
    static constexpr size_t SMEM = (size_t)NW * PER_WARP * sizeof(double);
};

// gauss_point_eval (en234gpu_kernels.cuh) with the element's node data (x, y, z, u1, u2, u3 per node) already staged in
// shared memory: same operations in the same order, so the results are bitwise those of the global-memory version.
template <int C>
__device__ __forceinline__ void jac_acc_sm(const double* nb, const double fp[3], const double fm[3], double* J) {
    const double xc = nb[6 * C], yc = nb[6 * C + 1], zc = nb[6 * C + 2];
    double dn[3];
    dshape<C>(fp, fm, dn);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        J[j] += xc * dn[j];
        J[3 + j] += yc * dn[j];
        J[6 + j] += zc * dn[j];
    }
}
template <int C>
__device__ __forceinline__ void grad_acc_sm(const double* nb, const double fp[3], const double fm[3], const double* Ji,
                                            double* g, double* H) {
    double dn[3], d[3];
    dshape<C>(fp, fm, dn);
#pragma unroll
    for (int j = 0; j < 3; ++j) d[j] = dn[0] * Ji[j] + dn[1] * Ji[3 + j] + dn[2] * Ji[6 + j];
    g[3 * C] = d[0]; g[3 * C + 1] = d[1]; g[3 * C + 2] = d[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double u = nb[6 * C + 3 + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) H[3 * i + j] += u * d[j];
    }
}
// the second half of gauss_point_eval_sm: dN/dx and grad u from a given inverse Jacobian (same operations, same order)
__device__ __forceinline__ void gauss_point_grad_sm(const double* nb, int gp, const Jac& jc, double* g, GpOut& o) {
    double fp[3], fm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double xi = ((gp >> j) & 1) ? EN234_GP : -EN234_GP;
        fp[j] = 1.0 + xi;
        fm[j] = 1.0 - xi;
    }
    o.det = jc.det;
#pragma unroll
    for (int q = 0; q < 9; ++q) o.H[q] = 0.0;
    grad_acc_sm<0>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<1>(nb, fp, fm, jc.Ji, g, o.H);
    grad_acc_sm<2>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<3>(nb, fp, fm, jc.Ji, g, o.H);
    grad_acc_sm<4>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<5>(nb, fp, fm, jc.Ji, g, o.H);
    grad_acc_sm<6>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<7>(nb, fp, fm, jc.Ji, g, o.H);
    g[24] = jc.det;   // w = 1 for the 2x2x2 rule
}
__device__ __forceinline__ void gauss_point_eval_sm(const double* nb, int gp, double* g, GpOut& o, Jac* jc_out = nullptr) {
    double fp[3], fm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double xi = ((gp >> j) & 1) ? EN234_GP : -EN234_GP;
        fp[j] = 1.0 + xi;
        fm[j] = 1.0 - xi;
    }
    double J[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) J[q] = 0.0;
    jac_acc_sm<0>(nb, fp, fm, J); jac_acc_sm<1>(nb, fp, fm, J); jac_acc_sm<2>(nb, fp, fm, J); jac_acc_sm<3>(nb, fp, fm, J);
    jac_acc_sm<4>(nb, fp, fm, J); jac_acc_sm<5>(nb, fp, fm, J); jac_acc_sm<6>(nb, fp, fm, J); jac_acc_sm<7>(nb, fp, fm, J);
    Jac jc;
    invert3(J, jc);
    if (jc_out) *jc_out = jc;
    o.det = jc.det;
#pragma unroll
    for (int q = 0; q < 9; ++q) o.H[q] = 0.0;
    grad_acc_sm<0>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<1>(nb, fp, fm, jc.Ji, g, o.H);
    grad_acc_sm<2>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<3>(nb, fp, fm, jc.Ji, g, o.H);
    grad_acc_sm<4>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<5>(nb, fp, fm, jc.Ji, g, o.H);
    grad_acc_sm<6>(nb, fp, fm, jc.Ji, g, o.H); grad_acc_sm<7>(nb, fp, fm, jc.Ji, g, o.H);
    g[24] = jc.det;   // w = 1 for the 2x2x2 rule
}

// One warp per group of four elements.
//  stage    lane = 8*g + p : coordinates and DOFs of node p of element g, loaded one group AHEAD into registers (the
//           global-memory latency overlaps the arithmetic of the current group) and handed over through shared memory
//  phase A  lane = 8*g + p : Gauss point p of element g — Jacobian, dN/dx, grad u, stress -> shared memory
//  phase R  lane = 8*g + a : residual of local node a,  R_a = -sum_p w|J| sigma_p . grad N_a   (el_linelast_3Dbasic.f90:128)
//  phase B  lane = 16*h + 4*tA + tB (two rounds, elements 2*round + h): the 2x2 tile of 3x3 blocks
//           K_e[a][b] = sum_p w|J| (lam S + mu S^T + mu tr(S) I),  S = grad N_a (x) grad N_b
//           — the integrand of B^T D B (el_linelast_3Dbasic.f90:130-131) with the zeros of B and D removed; the tile keeps
//           36 accumulators in registers so every shared-memory operand is used three times.
// BLOCKS = false: residual only (matrix-free residual assembly, and the matrix-free operator itself: with
// Load = LoadMasked the "residual" of the displacement field x is -K_e x_e).
template <bool NEO, bool BLOCKS, class Load>
__global__ void __launch_bounds__(ElemCfg<NEO>::NW * 32, 2) k_element_blocks(ElemArgs A, Load ld) {
    constexpr int NW = ElemCfg<NEO>::NW, GS = ElemCfg<NEO>::GS, G_SIG = ElemCfg<NEO>::G_SIG;
    extern __shared__ double sm[];
    if (ld.stopped()) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NS = ElemCfg<NEO>::NSTRIDE;
    double* geom = sm + (size_t)warp * ElemCfg<NEO>::PER_WARP;
    double* nbuf = geom + 4 * 8 * GS;
    const Mesh& M = A.M;
    const int ngroups = (M.E + 3) >> 2;
    const Material mat0 = M.mats[0];
    bool bad = false;
    double nd[6];
    const int W = gridDim.x * NW;
    auto node_of = [&](int grp) -> int {                             // node (lane & 7) of element 4 grp + (lane >> 3)
        if (grp >= ngroups) return 0;
        return M.conn[(size_t)(lane & 7) * M.E + min(4 * grp + (lane >> 3), M.E - 1)];   // padding lanes repeat the last element
    };
    auto stage = [&](int n) {
        nd[0] = __ldg(M.x + n); nd[1] = __ldg(M.y + n); nd[2] = __ldg(M.z + n);
        nd[3] = ld(3 * (size_t)n); nd[4] = ld(3 * (size_t)n + 1); nd[5] = ld(3 * (size_t)n + 2);
    };
    int grp = blockIdx.x * NW + warp;
    stage(node_of(grp));
    int n_next = node_of(grp + W);                                   // node index fetched two groups ahead of its use
    for (; grp < ngroups; grp += W) {
        {
            double* dst = nbuf + (lane >> 3) * NS + (lane & 7) * 6;
#pragma unroll
            for (int q = 0; q < 6; ++q) dst[q] = nd[q];
        }
        __syncwarp();
        stage(n_next);
        n_next = node_of(grp + 2 * W);
        // scratch position of this lane's phase-R output, fetched now and used after phase A
        const int rpos = 4 * grp + (lane >> 3) < M.E ? M.adj_pos[8 * (size_t)(4 * grp + (lane >> 3)) + (lane & 7)] : 0;
        // ---------------- phase A
        {
            const int g = lane >> 3, p = lane & 7;
            const int e = min(4 * grp + g, M.E - 1);
            const bool valid = 4 * grp + g < M.E;
            const Material mat = M.n_mats == 1 ? mat0 : M.mats[M.mat[e]];
            double* gg = geom + (g * 8 + p) * GS;
            GpOut gp;
            gauss_point_eval_sm(nbuf + g * NS, p, gg, gp);
            if (valid && !(gp.det != 0.0)) bad = true;               // zero / NaN Jacobian (Element_Utilities.f90:108-112)
            double s[9];
            if (NEO && mat.kind == 1) {
                NeoPoint np;
                neo_eval(mat, gp.H, gp.det, np);
                neo_node_vectors(np, gg);
                if (valid && !(np.J > 0.0)) bad = true;              // inverted element: cut the step back
#pragma unroll
                for (int q = 0; q < 9; ++q) s[q] = np.P[q];
            } else {
                stress_linear(mat, gp.H, s);
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) gg[G_SIG + q] = s[q];
        }
        __syncwarp();
        // ---------------- phase R
        {
            const int g = lane >> 3, a = lane & 7;
            const int e = 4 * grp + g;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const double* gq = geom + (g * 8 + p) * GS;
                const double* s = gq + G_SIG;
                const double w = gq[G_W];
                const double d0 = gq[3 * a], d1 = gq[3 * a + 1], d2 = gq[3 * a + 2];
                r0 -= (s[0] * d0 + s[1] * d1 + s[2] * d2) * w;
                r1 -= (s[3] * d0 + s[4] * d1 + s[5] * d2) * w;
                r2 -= (s[6] * d0 + s[7] * d1 + s[8] * d2) * w;
            }
            if (e < M.E) {
                double* R = A.Re + 3 * (size_t)rpos;
                R[0] = r0; R[1] = r1; R[2] = r2;
            }
        }
        // ---------------- phase B
        if (BLOCKS) {
            const int h = lane >> 4, tA = (lane >> 2) & 3, tB = lane & 3;
#pragma unroll 1
            for (int round = 0; round < 2; ++round) {
                const int el = 2 * round + h;
                const int e = min(4 * grp + el, M.E - 1);
                const Material mat = M.n_mats == 1 ? mat0 : M.mats[M.mat[e]];
                const double* ge = geom + el * 8 * GS;
                double K[4][9];
                if (NEO && mat.kind == 1) {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int q = 0; q < 9; ++q) K[t][q] = 0.0;
#pragma unroll 1
                    for (int p = 0; p < 8; ++p) {
                        const double* gq = ge + p * GS;
                        neo_block_acc(gq, 2 * tA, 2 * tB, K[0]);
                        neo_block_acc(gq, 2 * tA, 2 * tB + 1, K[1]);
                        neo_block_acc(gq, 2 * tA + 1, 2 * tB, K[2]);
                        neo_block_acc(gq, 2 * tA + 1, 2 * tB + 1, K[3]);
                    }
                } else {
                    double S[4][9];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int q = 0; q < 9; ++q) S[t][q] = 0.0;
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const double* gq = ge + p * GS;
                        const double w = gq[G_W];
                        double av[6], bv[6];
#pragma unroll
                        for (int q = 0; q < 6; ++q) { av[q] = w * gq[6 * tA + q]; bv[q] = gq[6 * tB + q]; }
#pragma unroll
                        for (int sa = 0; sa < 2; ++sa)
#pragma unroll
                            for (int sb = 0; sb < 2; ++sb)
#pragma unroll
                                for (int r = 0; r < 3; ++r)
#pragma unroll
                                    for (int c = 0; c < 3; ++c) S[sa * 2 + sb][3 * r + c] += av[3 * sa + r] * bv[3 * sb + c];
                    }
                    const double lam = mat.d12, mu = mat.d44;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const double tr = mu * (S[t][0] + S[t][4] + S[t][8]);
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                K[t][3 * r + c] = lam * S[t][3 * r + c] + mu * S[t][3 * c + r] + (r == c ? tr : 0.0);
                    }
                }
                // elements 4grp+2round, +1 form one scratch pair: 36 coalesced 256-byte rows
                double2* dst = reinterpret_cast<double2*>(A.Ke + ((size_t)(2 * grp + round)) * 1152) + lane;
#pragma unroll
                for (int jp = 0; jp < 18; ++jp) {
                    const int j0 = 2 * jp, j1 = 2 * jp + 1;
                    __stcs(dst + jp * 32, make_double2(K[j0 / 9][j0 % 9], K[j1 / 9][j1 % 9]));
                }
            }
        }
        __syncwarp();
    }
    if (bad && A.S) atomicOr(&A.S->nanflag, 1);
}

// Residual-only element kernel for the linear-elastic elements (matrix-free residual assembly and the general-mesh
// matrix-free operator, where the "residual" of the field x is -K_e x_e).  Same staging as k_element_blocks; each lane
// keeps its Gauss point's dN/dx in registers, forms the 24 nodal force entries of that point and the eight points of an
// element are summed by a halving butterfly over the eight lanes (lane a ends with node a) — no geometry in shared
// memory, so three CTAs fit per SM.
#ifndef FORCES_OCC
#define FORCES_OCC 2
#endif
// GEOM = 0: the Gauss-point geometry (Jacobian, cofactor inverse) is evaluated from the node coordinates;
// GEOM = 1: the same, and J^-1, w|J| are stored in A.geom;  GEOM = 2: they are read from A.geom (streamed one group ahead) —
// the node coordinates are then not touched at all and 40 % of the arithmetic of an element application is gone; the
// results are bitwise those of GEOM = 0 (the stored values are the ones it computes, the remaining operations are the same).
template <class Load, int GEOM = 0>
__global__ void __launch_bounds__(256, FORCES_OCC) k_element_forces(ElemArgs A, Load ld) {
    constexpr int NW = 8, NS = ElemCfg<false>::NSTRIDE;
    __shared__ double nbuf_all[NW * 4 * NS];
    if (ld.stopped()) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, p = lane & 7;
    double* nbuf = nbuf_all + warp * 4 * NS;
    const Mesh& M = A.M;
    const int ngroups = (M.E + 3) >> 2, W = gridDim.x * NW;
    const Material mat0 = M.mats[0];
    bool bad = false;
    double nd[6];
    auto node_of = [&](int grp) -> int {
        if (grp >= ngroups) return 0;
        return M.conn[(size_t)p * M.E + min(4 * grp + g, M.E - 1)];
    };
    auto stage = [&](int n) {
        if (GEOM != 2) { nd[0] = __ldg(M.x + n); nd[1] = __ldg(M.y + n); nd[2] = __ldg(M.z + n); }
        nd[3] = ld(3 * (size_t)n); nd[4] = ld(3 * (size_t)n + 1); nd[5] = ld(3 * (size_t)n + 2);
    };
    Jac jn;                                                           // GEOM = 2: geometry of the NEXT group, in flight
    auto fetch_geom = [&](int gr) {
        if (GEOM == 2 && gr < ngroups) {
            const double* src = A.geom + (size_t)gr * 320 + lane;
#pragma unroll
            for (int q = 0; q < 9; ++q) jn.Ji[q] = __ldcs(src + 32 * q);
            jn.det = __ldcs(src + 32 * 9);
        }
    };
    int grp = blockIdx.x * NW + warp;
    stage(node_of(grp));
    fetch_geom(grp);
    int n_next = node_of(grp + W);
    for (; grp < ngroups; grp += W) {
        {
            double* dst = nbuf + g * NS + p * 6;
#pragma unroll
            for (int q = (GEOM == 2 ? 3 : 0); q < 6; ++q) dst[q] = nd[q];
        }
        __syncwarp();
        Jac jc = jn;
        stage(n_next);
        fetch_geom(grp + W);
        n_next = node_of(grp + 2 * W);
        const int e = 4 * grp + g;
        const bool valid = e < M.E;
        const int pos = valid ? M.adj_pos[8 * (size_t)e + p] : 0;     // fetched now, used after the arithmetic
        const Material mat = M.n_mats == 1 ? mat0 : M.mats[M.mat[valid ? e : M.E - 1]];
        double gg[GEOM_STRIDE];
        GpOut gp;
        if (GEOM == 2) gauss_point_grad_sm(nbuf + g * NS, p, jc, gg, gp);
        else gauss_point_eval_sm(nbuf + g * NS, p, gg, gp, &jc);
        if (GEOM == 1) {
            if (A.geom_stride != 0) {
                if (valid) {
                    double* dst = A.geom + (size_t)(10 * p) * A.geom_stride + e;
#pragma unroll
                    for (int q = 0; q < 9; ++q) dst[(size_t)q * A.geom_stride] = jc.Ji[q];
                    dst[(size_t)9 * A.geom_stride] = jc.det;
                }
            } else {
                double* dst = A.geom + (size_t)grp * 320 + lane;
#pragma unroll
                for (int q = 0; q < 9; ++q) dst[32 * q] = jc.Ji[q];
                dst[32 * 9] = jc.det;
            }
        }
        if (valid && !(gp.det != 0.0)) bad = true;
        double s[9];
        stress_linear(mat, gp.H, s);
#pragma unroll
        for (int q = 0; q < 9; ++q) s[q] *= -gp.det;                  // R = - sum_p w|J| sigma . grad N  (w = 1)
        double f[24];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            f[3 * a] = s[0] * gg[3 * a] + s[1] * gg[3 * a + 1] + s[2] * gg[3 * a + 2];
            f[3 * a + 1] = s[3] * gg[3 * a] + s[4] * gg[3 * a + 1] + s[5] * gg[3 * a + 2];
            f[3 * a + 2] = s[6] * gg[3 * a] + s[7] * gg[3 * a + 1] + s[8] * gg[3 * a + 2];
        }
        // halving butterfly over the 8 Gauss-point lanes: lane a ends with the 3 components of node a
        double h[12], q6[6], r3[3];
        {
            const bool up = p & 4;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const double send = up ? f[j] : f[12 + j], keep = up ? f[12 + j] : f[j];
                h[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        {
            const bool up = p & 2;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double send = up ? h[j] : h[6 + j], keep = up ? h[6 + j] : h[j];
                q6[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
        {
            const bool up = p & 1;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double send = up ? q6[j] : q6[3 + j], keep = up ? q6[3 + j] : q6[j];
                r3[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        }
        if (valid) {
            double* R = A.Re + 3 * (size_t)pos;
            R[0] = r3[0]; R[1] = r3[1]; R[2] = r3[2];
        }
        __syncwarp();
    }
    if (bad && A.S) atomicOr(&A.S->nanflag, 1);
}

// Residual-only element kernel on the CACHED Gauss-point geometry (J^-1 = dxi/dx and w|J| per point, written once by
// k_element_forces<Load, 1>): the same lane mapping and butterfly as k_element_forces, but the shape-function gradients in
// physical space are never formed.  With d_a = dN_a/dxi at the point (three products of (1 +- xi_j), sign applied last):
//   grad u = G J^-1,   G = sum_a u_a (x) d_a                 (72 + 27 FMA instead of 72 + 72)
//   f_a    = T d_a,    T = -(w|J| sigma) J^-T                 (27 + 72 FMA instead of 72)
// 24 registers of dN/dx disappear, which is what lets three CTAs share an SM (the kernel is latency-bound: ncu shows the
// FP64 pipe 39 % busy at 16 warps per SM).  The element sums are the same numbers in a different association, so results
// agree with k_element_forces to rounding (1e-15 relative), not bitwise.
template <int C>
__device__ __forceinline__ void xi_grad_acc(const double* nb, const double fp[3], const double fm[3], double* G) {
    double dn[3];
    dshape<C>(fp, fm, dn);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double u = nb[6 * C + 3 + i];
#pragma unroll
        for (int k = 0; k < 3; ++k) G[3 * i + k] += u * dn[k];
    }
}
// dN/dxi of node C4 + 4 * upper (the two nodes differ in the xi_3 factor and the sign of the third component only); same
// operation order as dshape<C>
template <int C4>
__device__ __forceinline__ void dshape_half(const double fp[3], const double fm[3], bool upper, double dn[3]) {
    const double f0 = sgn(C4, 0) > 0 ? fp[0] : fm[0];
    const double f1 = sgn(C4, 1) > 0 ? fp[1] : fm[1];
    const double f2 = upper ? fp[2] : fm[2];
    dn[0] = sgn(C4, 0) * (f1 * f2 * 0.125);
    dn[1] = sgn(C4, 1) * (f0 * f2 * 0.125);
    const double d2 = f0 * f1 * 0.125;
    dn[2] = upper ? d2 : -d2;
}
// forces of the four nodes C4 + 4 * upper: f[3 * C4 + i] = sum_k T[i][k] dN/dxi_k
template <int C4>
__device__ __forceinline__ void xi_force_half(const double fp[3], const double fm[3], bool upper, const double* T, double* f) {
    double dn[3];
    dshape_half<C4>(fp, fm, upper, dn);
#pragma unroll
    for (int i = 0; i < 3; ++i) f[3 * C4 + i] = T[3 * i] * dn[0] + T[3 * i + 1] * dn[1] + T[3 * i + 2] * dn[2];
}
template <class Load, int OCC = 2>
__global__ void __launch_bounds__(256, OCC) k_element_forces_cached(ElemArgs A, Load ld) {
    constexpr int NW = 8, NS = ElemCfg<false>::NSTRIDE;
    __shared__ double nbuf_all[NW * 4 * NS];
    if (ld.stopped()) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, p = lane & 7;
    double* nbuf = nbuf_all + warp * 4 * NS;
    const Mesh& M = A.M;
    const int ngroups = (M.E + 3) >> 2, W = gridDim.x * NW;
    const Material mat0 = M.mats[0];
    double fp[3], fm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double xi = ((p >> j) & 1) ? EN234_GP : -EN234_GP;
        fp[j] = 1.0 + xi;
        fm[j] = 1.0 - xi;
    }
    double nd[3];
    auto node_of = [&](int grp) -> int {
        if (grp >= ngroups) return 0;
        return M.conn[(size_t)p * M.E + min(4 * grp + g, M.E - 1)];
    };
    auto stage = [&](int n) { nd[0] = ld(3 * (size_t)n); nd[1] = ld(3 * (size_t)n + 1); nd[2] = ld(3 * (size_t)n + 2); };
    int grp = blockIdx.x * NW + warp;
    stage(node_of(grp));
    int n_next = node_of(grp + W);
    for (; grp < ngroups; grp += W) {
        {
            double* dst = nbuf + g * NS + p * 6 + 3;
            dst[0] = nd[0]; dst[1] = nd[1]; dst[2] = nd[2];
        }
        __syncwarp();
        // loads of this iteration: the group's cached geometry (used after G below), the next group's node values
        double Ji[9], det;
        {
            const double* src = A.geom + (size_t)grp * 320 + lane;
#pragma unroll
            for (int q = 0; q < 9; ++q) Ji[q] = __ldcs(src + 32 * q);
            det = __ldcs(src + 32 * 9);
        }
        stage(n_next);
        n_next = node_of(grp + 2 * W);
        const int e = 4 * grp + g;
        const bool valid = e < M.E;
        const int pos = valid ? M.adj_pos[8 * (size_t)e + p] : 0;     // fetched now, used after the arithmetic
        const Material mat = M.n_mats == 1 ? mat0 : M.mats[M.mat[valid ? e : M.E - 1]];
        const double* nb = nbuf + g * NS;
        double G[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) G[q] = 0.0;
        xi_grad_acc<0>(nb, fp, fm, G); xi_grad_acc<1>(nb, fp, fm, G); xi_grad_acc<2>(nb, fp, fm, G); xi_grad_acc<3>(nb, fp, fm, G);
        xi_grad_acc<4>(nb, fp, fm, G); xi_grad_acc<5>(nb, fp, fm, G); xi_grad_acc<6>(nb, fp, fm, G); xi_grad_acc<7>(nb, fp, fm, G);
        double T[9];
        {
            double H[9], s[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) H[3 * i + j] = G[3 * i] * Ji[j] + G[3 * i + 1] * Ji[3 + j] + G[3 * i + 2] * Ji[6 + j];
            stress_linear(mat, H, s);
#pragma unroll
            for (int q = 0; q < 9; ++q) s[q] *= -det;                 // R = - sum_p w|J| sigma . grad N  (w = 1)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) T[3 * i + k] = s[3 * i] * Ji[3 * k] + s[3 * i + 1] * Ji[3 * k + 1] + s[3 * i + 2] * Ji[3 * k + 2];
        }
        // halving butterfly over the 8 Gauss-point lanes (lane a ends with the 3 components of node a); the half that is
        // sent away is computed first, the half that is kept while the shuffles are in flight
        double h[12], q6[6], r3[3];
        {
            const bool up = p & 4;                                    // this lane keeps nodes 4..7 when up, else 0..3
            double snd[12];
            xi_force_half<0>(fp, fm, !up, T, snd); xi_force_half<1>(fp, fm, !up, T, snd);
            xi_force_half<2>(fp, fm, !up, T, snd); xi_force_half<3>(fp, fm, !up, T, snd);
#pragma unroll
            for (int j = 0; j < 12; ++j) snd[j] = __shfl_xor_sync(0xffffffffu, snd[j], 4);
            xi_force_half<0>(fp, fm, up, T, h); xi_force_half<1>(fp, fm, up, T, h);
            xi_force_half<2>(fp, fm, up, T, h); xi_force_half<3>(fp, fm, up, T, h);
#pragma unroll
            for (int j = 0; j < 12; ++j) h[j] += snd[j];
        }
        {
            const bool up = p & 2;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double send = up ? h[j] : h[6 + j], keep = up ? h[6 + j] : h[j];
                q6[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
        {
            const bool up = p & 1;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double send = up ? q6[j] : q6[3 + j], keep = up ? q6[3 + j] : q6[j];
                r3[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        }
        if (valid) {
            double* R = A.Re + 3 * (size_t)pos;
            R[0] = r3[0]; R[1] = r3[1]; R[2] = r3[2];
        }
        __syncwarp();
    }
}

// Thread-per-element form of the residual-only kernel on the cached geometry (the matrix-free operator's hot kernel).
// ncu on the warp-cooperative forms above (k_element_forces, k_element_forces_cached): 750 warp instructions per group of
// four elements of which 288 are FP64 — issue slots, not the FP64 pipe (39 %) or HBM (36 %), bound them: the lane = (element,
// Gauss point) mapping pays for its butterflies, 64-bit selects and shared-memory hand-over.  Here one thread owns one
// element: the Gauss-point loop is unrolled, so dN_a/dxi are literal constants (no dshape arithmetic, no selects), the 24
// nodal values and the 24 force accumulators stay in registers, there is no shared memory, no shuffle and no divergence;
// what is left per point is G (72 FMA), H = G J^-1 (27), the stress, T (27) and f += T d (72).
// Geometry cache layout for this kernel: geomT[(10 * point + q) * Epad + e] (coalesced over e), written by
// k_element_forces<Load, 1> when ElemArgs::geom_stride != 0.
template <int P, int C>
__device__ __forceinline__ void tpe_grad(const double (&u)[24], double* G) {
    constexpr double A = 1.0 + EN234_GP, B = 1.0 - EN234_GP;
    constexpr double f0 = (sgn(C, 0) > 0) == (((P >> 0) & 1) != 0) ? A : B;
    constexpr double f1 = (sgn(C, 1) > 0) == (((P >> 1) & 1) != 0) ? A : B;
    constexpr double f2 = (sgn(C, 2) > 0) == (((P >> 2) & 1) != 0) ? A : B;
    constexpr double d0 = sgn(C, 0) * (f1 * f2 * 0.125), d1 = sgn(C, 1) * (f0 * f2 * 0.125), d2 = sgn(C, 2) * (f0 * f1 * 0.125);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        G[3 * i] += u[3 * C + i] * d0;
        G[3 * i + 1] += u[3 * C + i] * d1;
        G[3 * i + 2] += u[3 * C + i] * d2;
    }
}
template <int P, int C>
__device__ __forceinline__ void tpe_force(const double* T, double (&f)[24]) {
    constexpr double A = 1.0 + EN234_GP, B = 1.0 - EN234_GP;
    constexpr double f0 = (sgn(C, 0) > 0) == (((P >> 0) & 1) != 0) ? A : B;
    constexpr double f1 = (sgn(C, 1) > 0) == (((P >> 1) & 1) != 0) ? A : B;
    constexpr double f2 = (sgn(C, 2) > 0) == (((P >> 2) & 1) != 0) ? A : B;
    constexpr double d0 = sgn(C, 0) * (f1 * f2 * 0.125), d1 = sgn(C, 1) * (f0 * f2 * 0.125), d2 = sgn(C, 2) * (f0 * f1 * 0.125);
#pragma unroll
    for (int i = 0; i < 3; ++i) f[3 * C + i] += T[3 * i] * d0 + T[3 * i + 1] * d1 + T[3 * i + 2] * d2;
}
template <int P>
__device__ __forceinline__ void tpe_point(const double* __restrict__ geomT, size_t stride, const Material& mat,
                                          const double (&u)[24], double (&f)[24]) {
    double Ji[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) Ji[q] = __ldcs(geomT + (size_t)(10 * P + q) * stride);
    const double det = __ldcs(geomT + (size_t)(10 * P + 9) * stride);
    double G[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) G[q] = 0.0;
    tpe_grad<P, 0>(u, G); tpe_grad<P, 1>(u, G); tpe_grad<P, 2>(u, G); tpe_grad<P, 3>(u, G);
    tpe_grad<P, 4>(u, G); tpe_grad<P, 5>(u, G); tpe_grad<P, 6>(u, G); tpe_grad<P, 7>(u, G);
    double H[9], s[9], T[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) H[3 * i + j] = G[3 * i] * Ji[j] + G[3 * i + 1] * Ji[3 + j] + G[3 * i + 2] * Ji[6 + j];
    stress_linear(mat, H, s);
#pragma unroll
    for (int q = 0; q < 9; ++q) s[q] *= -det;                         // R = - sum_p w|J| sigma . grad N  (w = 1)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) T[3 * i + k] = s[3 * i] * Ji[3 * k] + s[3 * i + 1] * Ji[3 * k + 1] + s[3 * i + 2] * Ji[3 * k + 2];
    tpe_force<P, 0>(T, f); tpe_force<P, 1>(T, f); tpe_force<P, 2>(T, f); tpe_force<P, 3>(T, f);
    tpe_force<P, 4>(T, f); tpe_force<P, 5>(T, f); tpe_force<P, 6>(T, f); tpe_force<P, 7>(T, f);
}
constexpr int TPE_THREADS = 128;
#ifndef TPE_OCC
#define TPE_OCC 2
#endif
template <class Load>
__global__ void __launch_bounds__(TPE_THREADS, TPE_OCC) k_element_forces_tpe(ElemArgs A, Load ld) {
    if (ld.stopped()) return;
    const Mesh& M = A.M;
    const Material mat0 = M.mats[0];
    for (int e = blockIdx.x * TPE_THREADS + threadIdx.x; e < M.E; e += gridDim.x * TPE_THREADS) {
        const Material mat = M.n_mats == 1 ? mat0 : M.mats[M.mat[e]];
        double u[24], f[24];
        int pos[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int n = M.conn[(size_t)a * M.E + e];
            pos[a] = M.adj_pos[8 * (size_t)e + a];
            u[3 * a] = ld(3 * (size_t)n); u[3 * a + 1] = ld(3 * (size_t)n + 1); u[3 * a + 2] = ld(3 * (size_t)n + 2);
        }
#pragma unroll
        for (int q = 0; q < 24; ++q) f[q] = 0.0;
        const double* gT = A.geom + e;
        const size_t st = A.geom_stride;
        tpe_point<0>(gT, st, mat, u, f); tpe_point<1>(gT, st, mat, u, f); tpe_point<2>(gT, st, mat, u, f); tpe_point<3>(gT, st, mat, u, f);
        tpe_point<4>(gT, st, mat, u, f); tpe_point<5>(gT, st, mat, u, f); tpe_point<6>(gT, st, mat, u, f); tpe_point<7>(gT, st, mat, u, f);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            double* R = A.Re + 3 * (size_t)pos[a];
            R[0] = f[3 * a]; R[1] = f[3 * a + 1]; R[2] = f[3 * a + 2];
        }
    }
}

#ifndef EN234_GATHER_MINB
#define EN234_GATHER_MINB 2
#endif
struct GatherArgs {
    Mesh M;
    const double* Ke;
    const double* Re;
    const double* felem;    // element-level external loads added to the residual (may be null)
    double* val;            // block-row values, layout per block row i: [r][block][c]  (= scalar CSR of rows 3i..3i+2)
    double* rhs;            // pre-BC right-hand side (element residual sum + felem)
    double* rforce;         // -rhs                                        (solver_direct.f90:417-421)
    double* partials;       // per-CTA sum of rhs^2                        (:423)
    CgScalars* S;           // nanflag                                     (:402-415)
    double* diag;           // assembled diagonal per DOF (Jacobi preconditioner input, cg_iteration_loop :817-820)
    const unsigned* inv;    // per block (row, slot): nibble k = local node b of adjacency entry k on that column, 15 = none
    int assemble_matrix;
};

// One warp per node row, ONE LANE PER 3x3 BLOCK of the row (a hex8 row has <= 27 blocks).  `inv` holds, for every block
// (row i, slot s), one nibble per adjacency entry k < 8 of the row: the local node b of that element whose column is s,
// or 15 when the element does not touch the column.  A lane therefore knows statically-indexed where its up to 8
// contributions live in the element scratch: 9 accumulators in registers, contributions added in ascending element
// order (the reference's serial order), no shared memory, no atomics; all loads of a row are independent.
// Row pointers are fetched two rows ahead and adjacency / inverse-map words one row ahead, so one row costs one
// memory latency.  Rows with more than 8 adjacent elements or more than 32 blocks take the generic path.
__device__ __forceinline__ void gather_row_generic(const GatherArgs& A, int i, int k0, int k1, int rb, int nb, int lane,
                                                   double* accs, double& R) {
    const Mesh& M = A.M;
    const int L = 3 * nb;
    double* acc = (nb <= MAXNB) ? accs : A.val + 9 * (size_t)rb;
    if (A.assemble_matrix) {
        for (int q = lane; q < 9 * nb; q += 32) acc[q] = 0.0;
    }
    __syncwarp();
    for (int k = k0; k < k1; ++k) {
        const int ea = M.adj[k];
        const int e = ea >> 3, a = ea & 7;
        if (lane < 3) R += A.Re[3 * (size_t)k + lane];
        if (A.assemble_matrix) {
            for (int q = lane; q < 72; q += 32) {
                const int b = q / 9, rc = q % 9, j = ((a & 1) * 2 + (b & 1)) * 9 + rc;
                const double v = A.Ke[(size_t)(e >> 1) * 1152 + (j >> 1) * 64 + ((e & 1) * 16 + (a >> 1) * 4 + (b >> 1)) * 2 + (j & 1)];
                acc[(rc / 3) * L + 3 * (int)M.slot[(size_t)k * 8 + b] + rc % 3] += v;
            }
        }
        __syncwarp();
    }
    if (A.assemble_matrix) {
        const int dslot = k0 < k1 ? M.slot[(size_t)k0 * 8 + (M.adj[k0] & 7)] : 0;
        if (lane < 3) A.diag[3 * (size_t)i + lane] = acc[lane * L + 3 * dslot + lane];
        if (nb <= MAXNB) {
            double* out = A.val + 9 * (size_t)rb;
            for (int q = lane; q < 9 * nb; q += 32) out[q] = acc[q];
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(THREADS, EN234_GATHER_MINB) k_gather_rows(GatherArgs A) {
    __shared__ double accs[WARPS * ACC_PER_WARP];
    __shared__ double red[WARPS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Mesh& M = A.M;
    const int W = gridDim.x * WARPS;
    double norm2 = 0.0;
    bool bad = false;
    int i = blockIdx.x * WARPS + warp;
    // software pipeline: (k0, k1, rb, re) of rows i and i + W, adjacency entry / inverse-map word of row i
    int k0 = 0, k1 = 0, rb = 0, re = 0, k0n = 0, k1n = 0, rbn = 0, ren = 0, my_ea = -1;
    unsigned my_inv = 0xffffffffu;
    if (i < M.N) { k0 = M.adj_ptr[i]; k1 = M.adj_ptr[i + 1]; rb = M.browptr[i]; re = M.browptr[i + 1]; }
    if (i + W < M.N) { k0n = M.adj_ptr[i + W]; k1n = M.adj_ptr[i + W + 1]; rbn = M.browptr[i + W]; ren = M.browptr[i + W + 1]; }
    if (lane < 8 && k0 + lane < k1) my_ea = M.adj[k0 + lane];
    if (A.assemble_matrix && rb + lane < re) my_inv = A.inv[rb + lane];
    for (; i < M.N; i += W) {
        const int nb = re - rb, cnt = k1 - k0, L = 3 * nb;
        double R = 0.0;
        // prefetch for the next row / the row after it
        int ea_n = -1, k0f = 0, k1f = 0, rbf = 0, ref = 0;
        unsigned inv_n = 0xffffffffu;
        if (lane < 8 && k0n + lane < k1n) ea_n = M.adj[k0n + lane];
        if (A.assemble_matrix && rbn + lane < ren) inv_n = A.inv[rbn + lane];
        if (i + 2 * W < M.N) {
            k0f = M.adj_ptr[i + 2 * W]; k1f = M.adj_ptr[i + 2 * W + 1]; rbf = M.browptr[i + 2 * W]; ref = M.browptr[i + 2 * W + 1];
        }
        if (cnt <= 8 && nb <= 32) {
            double acc[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q] = 0.0;
            // two halves of four contributions: all loads of a half are issued before its (ordered) adds
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                double v[4][9], rv[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = 4 * half + u;
                    const int ea = __shfl_sync(0xffffffffu, my_ea, k);
                    const int e = ea >> 3, a = ea & 7;
                    const int b = (my_inv >> (4 * k)) & 15;
                    ok[u] = A.assemble_matrix && ea >= 0 && b != 15;
                    rv[u] = 0.0;
                    if (lane < 3 && ea >= 0) rv[u] = A.Re[3 * (size_t)(k0 + k) + lane];
                    if (ok[u]) {
                        const double* src = A.Ke + (size_t)(e >> 1) * 1152 + ((e & 1) * 16 + (a >> 1) * 4 + (b >> 1)) * 2;
                        const int j0 = ((a & 1) * 2 + (b & 1)) * 9;
#pragma unroll
                        for (int rc = 0; rc < 9; ++rc) {
                            const int j = j0 + rc;
                            // default caching, not evict-first: the rest of each 64-byte chunk is wanted by the
                            // neighbouring lanes' later loads (with __ldcs the scratch came from HBM 1.45 times)
                            v[u][rc] = __ldg(src + (j >> 1) * 64 + (j & 1));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    R += rv[u];
                    if (ok[u]) {
#pragma unroll
                        for (int rc = 0; rc < 9; ++rc) acc[rc] += v[u][rc];
                    }
                }
                if (cnt <= 4) break;                                  // warp-uniform
            }
            if (A.assemble_matrix && lane < nb) {
                double* out = A.val + 9 * (size_t)rb + 3 * lane;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) out[r * L + c] = acc[3 * r + c];
                // the diagonal block is the one whose first contribution is (e, a) -> b == a
                const int ea0 = __shfl_sync(__activemask(), my_ea, 0);
                if (cnt > 0 && (int)(my_inv & 15) == (ea0 & 7)) {
                    double* dg = A.diag + 3 * (size_t)i;
                    dg[0] = acc[0]; dg[1] = acc[4]; dg[2] = acc[8];
                }
                if (cnt == 0 && lane == 0) { double* dg = A.diag + 3 * (size_t)i; dg[0] = dg[1] = dg[2] = 0.0; }
            }
        } else {
            gather_row_generic(A, i, k0, k1, rb, nb, lane, accs + warp * ACC_PER_WARP, R);
        }
        if (lane < 3) {
            double r = R;
            if (A.felem) r += A.felem[3 * (size_t)i + lane];
            A.rhs[3 * (size_t)i + lane] = r;
            A.rforce[3 * (size_t)i + lane] = -r;
            norm2 += r * r;
            if (r != r || r == r + 1.0) bad = true;              // NaN / Inf scan (solver_direct.f90:402-415)
        }
        k0 = k0n; k1 = k1n; rb = rbn; re = ren; my_ea = ea_n; my_inv = inv_n;
        k0n = k0f; k1n = k1f; rbn = rbf; ren = ref;
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
    const double t = block_sum<THREADS>(norm2, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
}

// Residual-only assembly (matrix-free solves): one thread per node sums the node's element residual entries in
// ascending element order; rhs, rforce = -rhs, ||rhs||^2 partials and the NaN scan as in k_gather_rows.
__global__ void __launch_bounds__(VEC_THREADS) k_gather_residual(GatherArgs A) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    const Mesh& M = A.M;
    double norm2 = 0.0;
    bool bad = false;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < M.N; i += gridDim.x * VEC_THREADS) {
        const int k0 = M.adj_ptr[i], k1 = M.adj_ptr[i + 1];
        double r[3] = {0.0, 0.0, 0.0};
        for (int k = k0; k < k1; ++k) {
            const double* f = A.Re + 3 * (size_t)k;
            r[0] += __ldcs(f); r[1] += __ldcs(f + 1); r[2] += __ldcs(f + 2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t d = 3 * (size_t)i + c;
            double v = r[c];
            if (A.felem) v += A.felem[d];
            A.rhs[d] = v;
            A.rforce[d] = -v;
            norm2 += v * v;
            if (v != v || v == v + 1.0) bad = true;
        }
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
    const double t = block_sum<VEC_THREADS>(norm2, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
}

}  // namespace en234k
