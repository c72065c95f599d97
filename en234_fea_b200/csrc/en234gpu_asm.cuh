// Assembly / boundary-condition kernels (see en234gpu_kernels.cuh for the reference citations).
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

struct AsmArgs {
    Mesh M;
    const double* ut;       // dof_total      (AoS, 3 per node)
    const double* du;       // dof_increment
    const double* felem;    // element-level external loads added to the residual (may be null)
    double* val;            // block-row values, layout per block row i: [r][block][c]  (= scalar CSR of rows 3i..3i+2)
    double* rhs;            // pre-BC right-hand side (element residual sum + felem)
    double* rforce;         // -rhs
    double* partials;       // per-CTA sum of rhs^2
    CgScalars* S;           // nanflag
    int assemble_matrix;    // 0: residual only (matrix-free solves)
};

// One warp per node row i.  The row's adjacency entries (element e, local node a) are processed four at a time:
// lane = 8*g + b.  Phase A: lane b evaluates Gauss point b of element e_g (Jacobian, dN/dx, grad u, stress) and
// leaves dN/dx, w|J| in shared memory.  Phase B: lane b forms the 3x3 block K_e[a][b] = sum_gp w|J| (lam S + mu S^T +
// mu tr(S) I), S = grad N_a (x) grad N_b  — the same integrand as B^T D B (el_linelast_3Dbasic.f90:130-131) with the
// zeros of B and D removed.  The four groups then add their blocks to the row accumulator one after the other, so every
// entry is summed in ascending element order like the reference's serial loop (solver_direct.f90:102): bitwise
// run-to-run deterministic, no atomics.
// NEO = true adds the compressible neo-Hookean element (material kind 1) with its larger per-point record; meshes
// without such elements run the NEO = false instantiation.
template <bool NEO>
struct AsmCfg {
    static constexpr int NW = NEO ? 4 : WARPS;                       // warps per CTA
    static constexpr int GS = NEO ? GEOM_STRIDE_NEO : GEOM_STRIDE;   // doubles per Gauss point record
    static constexpr int PER_WARP = 4 * 8 * GS + ACC_PER_WARP;
    static constexpr size_t SMEM = (size_t)NW * PER_WARP * sizeof(double);
};

template <bool NEO>
__global__ void __launch_bounds__(AsmCfg<NEO>::NW * 32, NEO ? 1 : 2) k_assemble_rows(AsmArgs A) {
    constexpr int NW = AsmCfg<NEO>::NW, GS = AsmCfg<NEO>::GS;
    extern __shared__ double sm[];
    __shared__ double red[NW + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, b = lane & 7;
    double* geom = sm + (size_t)warp * AsmCfg<NEO>::PER_WARP;
    double* accs = geom + 4 * 8 * GS;
    const Mesh& M = A.M;
    double norm2 = 0.0;
    bool bad = false;
    for (int i = blockIdx.x * NW + warp; i < M.N; i += gridDim.x * NW) {
        const int k0 = M.adj_ptr[i], k1 = M.adj_ptr[i + 1];
        const int rb = M.browptr[i], nb = M.browptr[i + 1] - rb, L = 3 * nb;
        double* acc = (nb <= MAXNB) ? accs : A.val + 9 * (size_t)rb;
        if (A.assemble_matrix) {
            for (int q = lane; q < 9 * nb; q += 32) acc[q] = 0.0;
        }
        __syncwarp();
        double R0 = 0.0, R1 = 0.0, R2 = 0.0;
        for (int kk = k0; kk < k1; kk += 4) {
            const int k = kk + g;
            const bool valid = k < k1;
            const int ea = M.adj[valid ? k : k0];
            const int e = ea >> 3, a = ea & 7;
            const Material mat = M.mats[M.mat[e]];
            double* gg = geom + (g * 8 + b) * GS;
            GpOut gp;
            gauss_point_eval(M, e, b, LoadSum{A.ut, A.du}, gg, gp);
            if (valid && !(gp.det != 0.0)) bad = true;          // zero / NaN Jacobian (Element_Utilities.f90:108-112)
            // ---- residual of node a: -sum_gp w|J| sigma . grad N_a      (el_linelast_3Dbasic.f90:128)
            double s[9];
            if (NEO && mat.kind == 1) {
                NeoPoint np;
                neo_eval(mat, gp.H, gp.det, np);
                neo_node_vectors(np, gg);
                if (valid && !(np.J > 0.0)) bad = true;          // inverted element: cut the step back
#pragma unroll
                for (int q = 0; q < 9; ++q) s[q] = np.P[q];
            } else {
                stress_linear(mat, gp.H, s);
            }
            __syncwarp();
            const double da0 = gg[3 * a], da1 = gg[3 * a + 1], da2 = gg[3 * a + 2];
            double r0 = -(s[0] * da0 + s[1] * da1 + s[2] * da2) * gp.det;
            double r1 = -(s[3] * da0 + s[4] * da1 + s[5] * da2) * gp.det;
            double r2 = -(s[6] * da0 + s[7] * da1 + s[8] * da2) * gp.det;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                r0 += __shfl_xor_sync(0xffffffffu, r0, o);
                r1 += __shfl_xor_sync(0xffffffffu, r1, o);
                r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            }
#pragma unroll
            for (int gs = 0; gs < 4; ++gs) {
                const double v0 = __shfl_sync(0xffffffffu, r0, gs * 8);
                const double v1 = __shfl_sync(0xffffffffu, r1, gs * 8);
                const double v2 = __shfl_sync(0xffffffffu, r2, gs * 8);
                if (kk + gs < k1) { R0 += v0; R1 += v1; R2 += v2; }
            }
            if (A.assemble_matrix) {
                // ---- block K_e[a][b]
                double K[9];
                const double* ge = geom + g * 8 * GS;
                if (NEO && mat.kind == 1) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) K[q] = 0.0;
#pragma unroll
                    for (int p = 0; p < 8; ++p) neo_block_acc(ge + p * GS, a, b, K);
                } else {
                    double S[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) S[q] = 0.0;
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const double* gq = ge + p * GS;
                        const double w = gq[G_W];
                        const double a0 = w * gq[3 * a], a1 = w * gq[3 * a + 1], a2 = w * gq[3 * a + 2];
                        const double b0 = gq[3 * b], b1 = gq[3 * b + 1], b2 = gq[3 * b + 2];
                        S[0] += a0 * b0; S[1] += a0 * b1; S[2] += a0 * b2;
                        S[3] += a1 * b0; S[4] += a1 * b1; S[5] += a1 * b2;
                        S[6] += a2 * b0; S[7] += a2 * b1; S[8] += a2 * b2;
                    }
                    const double lam = mat.d12, mu = mat.d44;
                    const double tr = mu * (S[0] + S[4] + S[8]);
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) K[3 * r + c] = lam * S[3 * r + c] + mu * S[3 * c + r] + (r == c ? tr : 0.0);
                }
                const int sl = valid ? M.slot[(size_t)k * 8 + b] : 0;
#pragma unroll
                for (int gs = 0; gs < 4; ++gs) {
                    if (g == gs && valid) {
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) acc[r * L + 3 * sl + c] += K[3 * r + c];
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
        }
        if (A.assemble_matrix && nb <= MAXNB) {
            double* dst = A.val + 9 * (size_t)rb;
            for (int q = lane; q < 9 * nb; q += 32) dst[q] = acc[q];
        }
        if (lane < 3) {
            double r = lane == 0 ? R0 : (lane == 1 ? R1 : R2);
            if (A.felem) r += A.felem[3 * (size_t)i + lane];
            A.rhs[3 * (size_t)i + lane] = r;
            A.rforce[3 * (size_t)i + lane] = -r;
            norm2 += r * r;
            if (r != r || r == r + 1.0) bad = true;              // NaN / Inf scan (solver_direct.f90:402-415)
        }
        __syncwarp();
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
    const double t = block_sum<NW * 32>(norm2, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
}

struct BcArgs {
    int N;
    const int* browptr;
    const int* bcol;
    double* val;
    double* rhs;                 // in: pre-BC rhs; out: post-BC rhs
    const double* fext;          // may be null
    const uint8_t* cflag;        // per DOF: 1 = prescribed
    const double* cval;          // per DOF: correction value
    double* minv;                // Jacobi preconditioner 1/diag (cg_iteration_loop :817-820)
    double* partials;            // per-CTA sum of rhs^2
    // partitioned runs: interface rows are partial sums, so the additive pieces are written out, halo-summed and
    // finalised by k_bc_finalize; the unit diagonal of a prescribed DOF is placed on the owner rank only
    double* corr_out = nullptr;
    double* diag_out = nullptr;
    const uint8_t* owned = nullptr;
    // row subset: only rows that are prescribed or couple to a prescribed DOF change (null = every row); the other
    // rows are finished by k_bc_rest from the diagonal the assembly left in `diag`
    const int* rows = nullptr;
    int n_rows = 0;
};

// fixdof semantics for all prescribed DOFs in one pass (solver_conjugate_gradient.f90:673-713):
// rhs_j -= K_jn c_n for free j; row and column n zeroed; diag_n = 1; rhs_n = c_n.  One warp per node row.
__global__ void __launch_bounds__(THREADS) k_apply_bc(BcArgs B) {
    __shared__ double red[WARPS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double norm2 = 0.0;
    const int count = B.rows ? B.n_rows : B.N;
    for (int ii = blockIdx.x * WARPS + warp; ii < count; ii += gridDim.x * WARPS) {
        const int i = B.rows ? B.rows[ii] : ii;
        const int rb = B.browptr[i], nb = B.browptr[i + 1] - rb, L = 3 * nb;
        double* v = B.val + 9 * (size_t)rb;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int row = 3 * i + r;
            const bool crow = B.cflag[row] != 0;
            double corr = 0.0, diag = 0.0;
            for (int q = lane; q < L; q += 32) {
                const int j = 3 * B.bcol[rb + q / 3] + q % 3;
                double a = v[r * L + q];
                const bool ccol = B.cflag[j] != 0;
                if (j == row) {
                    if (crow) { a = (B.owned == nullptr || B.owned[i]) ? 1.0 : 0.0; v[r * L + q] = a; }
                    diag = a;
                } else if (ccol) {
                    if (!crow) corr += a * B.cval[j];
                    v[r * L + q] = 0.0;
                } else if (crow) {
                    v[r * L + q] = 0.0;
                }
            }
            corr = warp_sum(corr);
            diag = warp_sum(diag);
            if (lane == 0 && B.corr_out) {
                B.corr_out[row] = corr;
                B.diag_out[row] = diag;
            } else if (lane == 0) {
                double rr;
                if (crow) rr = B.cval[row];
                else {
                    rr = B.rhs[row];
                    if (B.fext) rr += B.fext[row];
                    rr -= corr;
                }
                B.rhs[row] = rr;
                B.minv[row] = 1.0 / diag;
                norm2 += rr * rr;
            }
        }
    }
    const double t = block_sum<THREADS>(norm2, red);
    if (threadIdx.x == 0) B.partials[blockIdx.x] = t;
}

// rows outside the subset of k_apply_bc: rhs += f_ext, 1/diag from the assembled diagonal, ||rhs||^2 partials
__global__ void __launch_bounds__(VEC_THREADS) k_bc_rest(int n, const uint8_t* affected /*per node*/, double* rhs,
                                                          const double* fext, const double* diag, double* minv,
                                                          double* partials) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        if (affected[d / 3]) continue;
        double r = rhs[d];
        if (fext) { r += fext[d]; rhs[d] = r; }
        minv[d] = 1.0 / diag[d];
        s += r * r;
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// corrections for the device-resident path: c = value - (dof_total + dof_increment)  (solver_direct.f90:705-709)
__global__ void k_set_corrections(int n, const int* dof, const double* value, const double* ut, const double* du,
                                  uint8_t* cflag, double* cval, int values_are_corrections) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = dof[i];
    cflag[d] = 1;
    cval[d] = values_are_corrections ? value[i] : value[i] - (ut[d] + du[d]);
}

// final fixed-order sum of per-CTA partials into a scalar slot (single CTA)
__global__ void __launch_bounds__(VEC_THREADS) k_reduce_to(const double* part, int n, double* out) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    const double s = reduce_partials<VEC_THREADS>(part, n, red);
    if (threadIdx.x == 0) *out = s;
}

}  // namespace en234k
