// Two-node tie constraints (CONSTRAIN NODE PAIRS / TIE NODES of the reference's CONSTRAINTS block): u[dof1] = u[dof2].
// Reference: src/solver_direct.f90:292-328 — one Lagrange multiplier per constraint, appended to the equation system
//     rhs(row1) -= lambda, rhs(row2) += lambda, K(row1, lam) = +1, K(row2, lam) = -1,
//     rhs(lam) = (u2 + du2) - (u1 + du1) - lambda eps,  K(lam, lam) = eps = 1e-12 |diag| / neq,
// solved by the skyline LDU (an indefinite saddle-point system), multipliers updated by the solution (:967-969), norms of the
// Newton log taken over all equations including the multiplier rows (:423, :736, :971).
//
// B200 form: the saddle-point system is not formed.  Every tied group of DOFs is reduced to one unknown (its root) inside the
// Jacobi-PCG loop of the static step: p is expanded to the members before the SpMV, A p is summed back onto the root after it
// (k_tie_expand / k_tie_reduce around the unchanged k_spmv; the p.Ap partials of the SpMV are already those of the reduced
// operator), members are masked out of the vector updates by a zero preconditioner entry.  That reduced matrix T^T K T is
// symmetric positive definite, so the PCG kernels and their CUDA graph are reused as they are.  The multipliers are recovered
// after the solve from the equilibrium rows (lambda + dlambda = rhs - K du at a tied row), so the right-hand side, the reaction
// forces and the three norms of the Newton log are formed exactly as the reference forms them.  The compliance eps (1e-12 of
// the mean diagonal) enters the norms; its effect on the displacements (a gap of eps lambda) is dropped.
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

struct TieTables {
    const int* counts;      // [0] groups, [1] members (device: the row subset changes with the prescribed DOFs, the graph does not)
    const int* g_root;      // per group: root DOF
    const int* g_ptr;       // per group: members [g_ptr[g], g_ptr[g+1]) of m_dof (root excluded), ascending DOF order
    const uint8_t* g_fixed; // per group: the root is a prescribed DOF (members follow it, nothing is solved for)
    const int* m_dof;       // per member: DOF
    const int* m_root;      // per member: root DOF
    const uint8_t* m_fixed;
};

// rhs[d] += sum of (-lambda | +lambda) over the constraints at d, in constraint order; rforce = -rhs (solver_direct.f90:306, :321, :417-421)
__global__ void k_tie_rhs(int n_tdof, const int* tdof, const int* tptr, const int* tent, const double* lambda, double* rhs, double* rforce) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tdof) return;
    const int d = tdof[i];
    double v = rhs[d];
    for (int k = tptr[i]; k < tptr[i + 1]; ++k) {
        const int e = tent[k];
        v = (e & 1) ? v + lambda[e >> 1] : v - lambda[e >> 1];
    }
    rhs[d] = v;
    rforce[d] = -v;
}

// gap[nc] = (u2 + du2) - (u1 + du1), with the prescribed corrections added where a tied DOF is prescribed (cflag != null: the
// multiplier row after fixdof has moved the prescribed column to the right-hand side, solver_direct.f90:767)
__global__ void k_tie_gap(int n, const int* d1, const int* d2, const double* ut, const double* du, const uint8_t* cflag, const double* cval,
                          double* gap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = d1[i], b = d2[i];
    double wa = ut[a] + du[a], wb = ut[b] + du[b];
    if (cflag) {
        if (cflag[a]) wa += cval[a];
        if (cflag[b]) wb += cval[b];
    }
    gap[i] = wb - wa;
}

__global__ void __launch_bounds__(VEC_THREADS) k_sumsq(int n, const double* x, double* partials) {
    __shared__ double sm[VEC_THREADS / 32];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += x[i] * x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < VEC_THREADS / 32; ++w) t += sm[w];
        partials[blockIdx.x] = t;
    }
}

// known part of the correction of a member: d0[m] = closing of the current gap to its root (+ the root's prescribed correction)
__global__ void k_tie_d0(TieTables T, const double* ut, const double* du, const double* cval, double* d0, double* mgap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T.counts[1]) return;
    const int m = T.m_dof[i], r = T.m_root[i];
    const double g = (ut[r] + du[r]) - (ut[m] + du[m]);
    mgap[i] = g;
    d0[m] = T.m_fixed[i] ? g + cval[r] : g;
}

__global__ void k_vec_sub(int n, const double* y, double* x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] -= y[i];
}

// right-hand side and Jacobi diagonal of the reduced system: member rows are added onto the root row (ascending DOF order),
// members are masked (rhs = 0, minv = 0)
__global__ void k_tie_reduce_rhs(TieTables T, double* rhs, double* minv) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T.counts[0]) return;
    const int r = T.g_root[g];
    if (T.g_fixed[g]) {
        for (int k = T.g_ptr[g]; k < T.g_ptr[g + 1]; ++k) { rhs[T.m_dof[k]] = 0.0; minv[T.m_dof[k]] = 0.0; }
        return;
    }
    double s = rhs[r], d = 1.0 / minv[r];
    for (int k = T.g_ptr[g]; k < T.g_ptr[g + 1]; ++k) {
        const int m = T.m_dof[k];
        s += rhs[m];
        d += 1.0 / minv[m];
        rhs[m] = 0.0;
        minv[m] = 0.0;
    }
    rhs[r] = s;
    minv[r] = 1.0 / d;
}

// p on the members of the free groups = p on the root (columns of the members of a prescribed group stay out: p[m] = 0)
__global__ void k_tie_expand(TieTables T, double* p, const CgScalars* S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T.counts[1] || (S && S->stop)) return;
    if (!T.m_fixed[i]) p[T.m_dof[i]] = p[T.m_root[i]];
}

// y = A p: member rows onto the root (ascending DOF order), members cleared in y and p
__global__ void k_tie_reduce(TieTables T, double* y, double* p, const CgScalars* S) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T.counts[0] || (S && S->stop)) return;
    const int r = T.g_root[g];
    if (T.g_fixed[g]) {
        for (int k = T.g_ptr[g]; k < T.g_ptr[g + 1]; ++k) y[T.m_dof[k]] = 0.0;
        return;
    }
    double s = y[r];
    for (int k = T.g_ptr[g]; k < T.g_ptr[g + 1]; ++k) {
        const int m = T.m_dof[k];
        s += y[m];
        y[m] = 0.0;
        p[m] = 0.0;
    }
    y[r] = s;
}

// correction of the members: that of the root plus the closing of the gap
__global__ void k_tie_expand_sol(TieTables T, const double* mgap, double* sol) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T.counts[1]) return;
    sol[T.m_dof[i]] = sol[T.m_root[i]] + mgap[i];
}

// out[i] = rhs0[d] - y[d] at the tied DOFs: the sum of the multiplier corrections acting on row d
__global__ void k_tie_rows(int n_tdof, const int* tdof, const double* rhs0, const double* y, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tdof) out[i] = rhs0[tdof[i]] - y[tdof[i]];
}

}  // namespace en234k
