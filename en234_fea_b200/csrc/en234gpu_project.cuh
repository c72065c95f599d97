// State print on the device: least-squares projection of integration-point fields to the nodes
// (src/printing.f90:457-490).  Three steps, all on the mesh structures of the hot path:
//   k_project_rhs    f_k(n)  = sum over the node's elements, ascending element number, of int field_k N^n dV
//                    (assemble_field_projection, printing.f90:1386-1668; element parts el_linelast_3Dbasic.f90:266-413,
//                    abaqus_tecplot_interface.f90:130-180, continuum_elements.f90:1206-1351)
//   k_project_mass   M(n,.) = int N^n N^b dV with the 27-point rule (assemble_projection_mass_matrix, printing.f90:1214-1383)
//                    written as diag(m,m,m) blocks into the 3x3-block CSR of the stiffness, so that three fields are
//                    solved per pass by the PCG kernels of the static step (the reference factorises a skyline copy of M);
//                    also the row-sum lumped matrix (:1096-1212)
//   pack / unpack    three fields <-> one DOF-shaped vector
// One thread per node: a node's sums are formed by a single thread in the reference's element order, so the result is
// deterministic and needs no atomics.  Not a timed path (once per printed step).
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

constexpr int PROJ_MAXF = 24;
constexpr int PROJ_THREADS = 128;
constexpr int FIELD_ZERO = -1, FIELD_MISES = 6, FIELD_STRAIN0 = 7, FIELD_STATE0 = 100;

struct ProjArgs {
    Mesh M;
    const double* ut;       // dof_total
    const double* du;       // dof_increment
    const double* sv;       // updated element state variables (source 1)
    int nsp;                // state variables per integration point
    int source;             // 0: stress = D B (u + du) of the linear elastic elements 1001 / U1;  1: C3D8 state variables
    int nf;
    int code[PROJ_MAXF];    // 0..5 s11 s22 s33 s12 s13 s23 | 6 Mises | 7..12 strain | 100+k state variable k of the point | -1 zero
    double* F;              // out: F[k*N + n]
};

// N_a at the point with 1 +- xi_j in fp / fm (Element_Utilities.f90:908-915)
__device__ __forceinline__ double shape_a(int a, const double fp[3], const double fm[3]) {
    const bool px = (a == 1 || a == 2 || a == 5 || a == 6), py = (a == 2 || a == 3 || a == 6 || a == 7), pz = a >= 4;
    return (px ? fp[0] : fm[0]) * (py ? fp[1] : fm[1]) * (pz ? fp[2] : fm[2]) / 8.0;
}
__device__ __forceinline__ double jac_det(const Mesh& M, int e, const double fp[3], const double fm[3]) {
    double J[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) J[q] = 0.0;
    jac_acc<0>(M, e, fp, fm, J); jac_acc<1>(M, e, fp, fm, J); jac_acc<2>(M, e, fp, fm, J); jac_acc<3>(M, e, fp, fm, J);
    jac_acc<4>(M, e, fp, fm, J); jac_acc<5>(M, e, fp, fm, J); jac_acc<6>(M, e, fp, fm, J); jac_acc<7>(M, e, fp, fm, J);
    return J[0] * J[4] * J[8] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - J[2] * J[4] * J[6];
}

__global__ void __launch_bounds__(PROJ_THREADS) k_project_rhs(ProjArgs A) {
    const int n = blockIdx.x * PROJ_THREADS + threadIdx.x;
    const Mesh& M = A.M;
    if (n >= M.N) return;
    double acc[PROJ_MAXF];
#pragma unroll
    for (int k = 0; k < PROJ_MAXF; ++k) acc[k] = 0.0;
    const int p0 = M.adj_ptr[n], p1 = M.adj_ptr[n + 1];
    for (int p = p0; p < p1; ++p) {
        const int ea = M.adj[p], e = ea >> 3, a = ea & 7;
        double el[PROJ_MAXF];                                   // this element's int f_k N^a dV (nodal_fieldvariables(k,a))
#pragma unroll
        for (int k = 0; k < PROJ_MAXF; ++k) el[k] = 0.0;
        for (int gp = 0; gp < 8; ++gp) {
            double fp[3], fm[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double xi = ((gp >> j) & 1) ? EN234_GP : -EN234_GP;
                fp[j] = 1.0 + xi;
                fm[j] = 1.0 - xi;
            }
            double s[6], en[6], det;
            const double* svp = nullptr;
            if (A.source == 0) {
                double g[GEOM_STRIDE];
                GpOut o;
                gauss_point_eval(M, e, gp, LoadSum{A.ut, A.du}, g, o);
                det = o.det;
                const Material mat = M.mats[M.n_mats == 1 ? 0 : M.mat[e]];
                double s9[9];
                stress_linear(mat, o.H, s9);
                s[0] = s9[0]; s[1] = s9[4]; s[2] = s9[8]; s[3] = s9[1]; s[4] = s9[2]; s[5] = s9[5];
#pragma unroll
                for (int q = 0; q < 6; ++q) en[q] = 0.0;
            } else {
                det = jac_det(M, e, fp, fm);
                svp = A.sv + ((size_t)e * 8 + gp) * A.nsp;
#pragma unroll
                for (int q = 0; q < 6; ++q) { s[q] = svp[q]; en[q] = svp[6 + q]; }
            }
            const double Na = shape_a(a, fp, fm);
            double smises = 0.0;
            {
                const double pr = (s[0] + s[1] + s[2]) / 3.0;
                const double d0 = s[0] - pr, d1 = s[1] - pr, d2 = s[2] - pr;
                smises = sqrt((d0 * d0 + d1 * d1 + d2 * d2) + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5])) * sqrt(1.5);
            }
#pragma unroll
            for (int k = 0; k < PROJ_MAXF; ++k) {
                if (k < A.nf) {
                    const int c = A.code[k];
                    double v = 0.0;
                    if (c >= 0 && c < 6) v = s[c];
                    else if (c == FIELD_MISES) v = smises;
                    else if (c >= FIELD_STRAIN0 && c < FIELD_STRAIN0 + 6) v = en[c - FIELD_STRAIN0];
                    else if (c >= FIELD_STATE0 && svp != nullptr) v = svp[c - FIELD_STATE0];
                    el[k] = el[k] + v * Na * det;                // stress*N*determinant*w, w = 1
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PROJ_MAXF; ++k) acc[k] = acc[k] + el[k];
    }
#pragma unroll
    for (int k = 0; k < PROJ_MAXF; ++k)
        if (k < A.nf) A.F[(size_t)k * M.N + n] = acc[k];
}

struct MassArgs {
    Mesh M;
    double* val;            // block-CSR values, [r][block][c] per block row: receives diag(m, m, m) blocks
    double* minv;           // 1 / M(n,n) for the three DOF slots of node n (raw_diag: M(n,n) itself)
    double* lump;           // row-sum lumped projection matrix, lump[lump_stride * n] (+1, +2 filled too when the stride is 3)
    int raw_diag;           // partitioned runs: diagonal and lumped entries of interface nodes are partial sums, the caller
    int lump_stride;        // halo-sums them (DOF-shaped vectors) and inverts afterwards
};

__global__ void __launch_bounds__(PROJ_THREADS) k_project_mass(MassArgs A) {
    const int n = blockIdx.x * PROJ_THREADS + threadIdx.x;
    const Mesh& M = A.M;
    if (n >= M.N) return;
    // Element_Utilities.f90:651-656: single-precision abscissa literal, truncated weights
    const double x1[3] = {-0.7745966911315918, 0.0, 0.7745966911315918};
    const double w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
    const int rb = M.browptr[n], nb = M.browptr[n + 1] - rb, L = 3 * nb;
    double* row = A.val + 9 * (size_t)rb;
    for (int q = 0; q < 3 * L; ++q) row[q] = 0.0;
    double lump = 0.0;
    int self = 0;
    const int p0 = M.adj_ptr[n], p1 = M.adj_ptr[n + 1];
    for (int p = p0; p < p1; ++p) {
        const int ea = M.adj[p], e = ea >> 3, a = ea & 7;
        double m[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = 0.0;
        double lm = 0.0;
        for (int kk = 0; kk < 3; ++kk)
            for (int jj = 0; jj < 3; ++jj)
                for (int ii = 0; ii < 3; ++ii) {
                    const double xi[3] = {x1[ii], x1[jj], x1[kk]};
                    const double w = w1[ii] * w1[jj] * w1[kk];
                    double fp[3], fm[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) { fp[j] = 1.0 + xi[j]; fm[j] = 1.0 - xi[j]; }
                    const double det = jac_det(M, e, fp, fm);
                    double Nn[8], sumN = 0.0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) { Nn[b] = shape_a(b, fp, fm); sumN += Nn[b]; }
                    const double Na = shape_a(a, fp, fm);
#pragma unroll
                    for (int b = 0; b < 8; ++b) m[b] = m[b] + Nn[b] * Na * det * w;
                    lm = lm + Na * sumN * det * w;
                }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int blk = M.slot[(size_t)p * 8 + b];
            if (b == a) self = blk;
            row[3 * blk] += m[b];
            row[L + 3 * blk + 1] += m[b];
            row[2 * L + 3 * blk + 2] += m[b];
        }
        lump = lump + lm;
    }
    double d = nb > 0 ? row[3 * self] : 1.0;
    if (p0 == p1 && nb > 0) {                                   // node without elements: identity row, its fields stay zero
        d = 1.0;
        row[3 * self] = 1.0; row[L + 3 * self + 1] = 1.0; row[2 * L + 3 * self + 2] = 1.0;
        lump = 1.0;
    }
    const double dv = A.raw_diag ? d : 1.0 / d;
    A.minv[3 * (size_t)n] = dv;
    A.minv[3 * (size_t)n + 1] = dv;
    A.minv[3 * (size_t)n + 2] = dv;
    for (int c = 0; c < A.lump_stride; ++c) A.lump[(size_t)A.lump_stride * n + c] = lump;
}
__global__ void k_project_invert(int n, double* v) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) v[d] = 1.0 / v[d];
}

// rhs[3n + r] = F[(k0 + r)*N + n] for the fields that exist (zero otherwise)
__global__ void k_project_pack(int N, int nf, int k0, const double* F, double* rhs) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
#pragma unroll
    for (int r = 0; r < 3; ++r) rhs[3 * (size_t)n + r] = (k0 + r < nf) ? F[(size_t)(k0 + r) * N + n] : 0.0;
}
__global__ void k_project_unpack(int N, int nf, int k0, const double* sol, double* F) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        if (k0 + r < nf) F[(size_t)(k0 + r) * N + n] = sol[3 * (size_t)n + r];
}
__global__ void k_project_lumped(int N, int nf, const double* lump, int lump_stride, double* F) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double l = lump[(size_t)lump_stride * n];
    for (int k = 0; k < nf; ++k) F[(size_t)k * N + n] = F[(size_t)k * N + n] / l;
}

}  // namespace en234k
