// Explicit dynamic step (src/explicit_dynamic_step.f90:1-83) on the device: central differences with the row-sum lumped
// mass matrix.  The element loop of assemble_dynamic_force (:394-718) is the residual-only element kernel of the static
// path (k_element_forces), so one time step is
//   k_dyn_predict         du = dt (v + dt/2 a), f = 0                                                    (:38-40)
//   k_dyn_load  x loads   f += s_l(t + dt) * (unit nodal forces of load l)         (apply_dynamic_boundaryconditions :796-1063)
//   k_dyn_pdof            du = value(t + dt) - u   |   dt * value  on the prescribed DOFs                   (:1064-1110)
//   k_element_forces      element residuals of (u, du) -> node-major scratch
//   k_dyn_gather_update   rforce = f + sum_e R_e (ascending element order, as the reference's loop adds them);
//                         a = rforce / m, v += dt a, u += du, t += dt                                     (:66-70)
// with the time and every vector resident on the device: no host round trip inside a chunk of steps.
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

struct DynTables {                 // history tables (modules/Boundaryconditions.f90:4-20) on the device
    const int* hist_ptr;           // n_histories + 1
    const double* hist_t;
    const double* hist_v;
};

// interpolate_history_table (modules/Boundaryconditions.f90:161-221)
__device__ __forceinline__ double dyn_history(const DynTables& T, int h /*1-based*/, double time) {
    const int o = T.hist_ptr[h - 1], n = T.hist_ptr[h] - o;
    const double* t = T.hist_t + o;
    const double* v = T.hist_v + o;
    if (n == 1) return v[0];
    if (time <= t[0]) return v[0];
    if (time >= t[n - 1]) return v[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int k = (hi + lo + 2) / 2 - 1;
        if (t[k] > time) hi = k; else lo = k;
    }
    if (t[hi] == t[lo]) return 0.5 * (v[hi] + v[lo]);
    return ((t[hi] - time) * v[lo] + (time - t[lo]) * v[hi]) / (t[hi] - t[lo]);
}

__global__ void __launch_bounds__(VEC_THREADS) k_dyn_predict(int n, double dt, const double* __restrict__ v,
                                                              const double* __restrict__ a, double* __restrict__ du,
                                                              double* __restrict__ f) {
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        du[d] = dt * (v[d] + 0.5 * dt * a[d]);
        f[d] = 0.0;
    }
}

// one load: f[dof_k] += scale * val_k, scale = history(t + dt) or the constant given (entries of one load have distinct DOFs)
__global__ void k_dyn_load(int n_entries, const int* __restrict__ dof, const double* __restrict__ val, int history,
                           double constant, DynTables T, const double* __restrict__ time, double dt, double* __restrict__ f) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_entries) return;
    const double s = history > 0 ? dyn_history(T, history, time[0] + dt) : constant;
    f[dof[k]] = f[dof[k]] + s * val[k];
}

// prescribed DOFs (distinct): mode 0 du = value - u; mode 1 du = dt * value (rate, node set); mode 2 du = value (rate, single node)
__global__ void k_dyn_pdof(int n_pd, const int* __restrict__ dof, const int* __restrict__ history, const double* __restrict__ value,
                           const int* __restrict__ mode, DynTables T, const double* __restrict__ time, double dt,
                           const double* __restrict__ ut, double* __restrict__ du) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pd) return;
    const double v = history[k] > 0 ? dyn_history(T, history[k], time[0] + dt) : value[k];
    const int d = dof[k], md = mode[k];
    du[d] = md == 0 ? v - ut[d] : (md == 1 ? dt * v : v);
}

struct DynUpdateArgs {
    Mesh M;
    const double* Re;      // node-major element residual scratch (k_element_forces)
    const double* f;       // external forces of this step
    const double* mass;    // lumped mass per DOF
    double *a, *v, *ut, *rforce;
    const double* du;
    double* time;
    double dt;
    CgScalars* S;          // nanflag
};

__global__ void __launch_bounds__(VEC_THREADS) k_dyn_gather_update(DynUpdateArgs A) {
    const Mesh& M = A.M;
    bool bad = false;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < M.N; i += gridDim.x * VEC_THREADS) {
        const int k0 = M.adj_ptr[i], k1 = M.adj_ptr[i + 1];
        double r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = A.f[3 * (size_t)i + c];
        for (int k = k0; k < k1; ++k) {
            const double* e = A.Re + 3 * (size_t)k;
            r[0] += __ldcs(e); r[1] += __ldcs(e + 1); r[2] += __ldcs(e + 2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t d = 3 * (size_t)i + c;
            const double m = A.mass[d];
            const double acc = m != 0.0 ? r[c] / m : 0.0;           // a node without elements has no mass: it stays at rest
            A.rforce[d] = r[c];
            A.a[d] = acc;
            A.v[d] = A.v[d] + A.dt * acc;
            A.ut[d] = A.ut[d] + A.du[d];
            if (acc != acc || acc == acc + 1.0) bad = true;
        }
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) A.time[0] = A.time[0] + A.dt;
}

// Partitioned runs: the element forces of interface nodes are partial sums, so the gather (k_gather_residual), the halo sum and
// this update are separate launches.  r = internal forces (halo-summed), f = external forces (full-valued on every copy).
__global__ void __launch_bounds__(VEC_THREADS) k_dyn_update(int n, double dt, const double* __restrict__ f, const double* __restrict__ r,
                                                             const double* __restrict__ mass, double* __restrict__ a,
                                                             double* __restrict__ v, double* __restrict__ ut,
                                                             const double* __restrict__ du, double* __restrict__ rforce,
                                                             double* time, CgScalars* S) {
    bool bad = false;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        const double rf = f[d] + r[d];
        const double m = mass[d];
        const double acc = m != 0.0 ? rf / m : 0.0;
        rforce[d] = rf;
        a[d] = acc;
        v[d] = v[d] + dt * acc;
        ut[d] = ut[d] + du[d];
        if (acc != acc || acc == acc + 1.0) bad = true;
    }
    if (bad) atomicOr(&S->nanflag, 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) time[0] = time[0] + dt;
}

// assemble_lumped_mass (:85-392): per element the row sums of the 27-point consistent matrix (user_element.f90:380-410 =
// abaqus_vuel.for:164-208) times the element's density, added per node in ascending element order.  One thread per node.
__global__ void __launch_bounds__(128) k_dyn_lumped_mass(Mesh M, const double* __restrict__ density, double* __restrict__ mass) {
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= M.N) return;
    const double x1[3] = {-0.7745966911315918, 0.0, 0.7745966911315918};
    const double w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
    double total = 0.0;
    for (int p = M.adj_ptr[n]; p < M.adj_ptr[n + 1]; ++p) {
        const int ea = M.adj[p], e = ea >> 3, a = ea & 7;
        double em = 0.0;
        for (int kk = 0; kk < 3; ++kk)
            for (int jj = 0; jj < 3; ++jj)
                for (int ii = 0; ii < 3; ++ii) {
                    const double xi[3] = {x1[ii], x1[jj], x1[kk]};
                    const double w = w1[ii] * w1[jj] * w1[kk];
                    double fp[3], fm[3], J[9];
#pragma unroll
                    for (int j = 0; j < 3; ++j) { fp[j] = 1.0 + xi[j]; fm[j] = 1.0 - xi[j]; }
#pragma unroll
                    for (int q = 0; q < 9; ++q) J[q] = 0.0;
                    jac_acc<0>(M, e, fp, fm, J); jac_acc<1>(M, e, fp, fm, J); jac_acc<2>(M, e, fp, fm, J); jac_acc<3>(M, e, fp, fm, J);
                    jac_acc<4>(M, e, fp, fm, J); jac_acc<5>(M, e, fp, fm, J); jac_acc<6>(M, e, fp, fm, J); jac_acc<7>(M, e, fp, fm, J);
                    const double det = J[0] * J[4] * J[8] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8] + J[1] * J[5] * J[6] +
                                       J[2] * J[3] * J[7] - J[2] * J[4] * J[6];
                    double sumN = 0.0, Na = 0.0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const bool px = (b == 1 || b == 2 || b == 5 || b == 6), py = (b == 2 || b == 3 || b == 6 || b == 7), pz = b >= 4;
                        const double Nb = (px ? fp[0] : fm[0]) * (py ? fp[1] : fm[1]) * (pz ? fp[2] : fm[2]) / 8.0;
                        sumN += Nb;
                        if (b == a) Na = Nb;
                    }
                    em = em + Na * sumN * det * w;
                }
        total = total + em * density[e];
    }
    mass[3 * (size_t)n] = total;
    mass[3 * (size_t)n + 1] = total;
    mass[3 * (size_t)n + 2] = total;
}

}  // namespace en234k
