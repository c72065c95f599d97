// Jacobi-preconditioned BiCGSTAB for UNSYMMETRIC stiffness matrices (SOLVER, ..., UNSYMMETRIC: src/solver_direct.f90
// stores alow for these, :282, :952-958; the reference's CG solver cannot handle them).  Used by the built-in C3D8
// continuum elements.  Plain first version: vector kernels with the recurrence scalars passed
// by value, the three scalar reductions of an iteration are read back by the host (no CUDA graph) — the systems of the
// reference's shipped continuum-element inputs have ~1e4 equations.  Deterministic: per-CTA partials in a fixed order.
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

// partials of a.b and c.d
__global__ void __launch_bounds__(VEC_THREADS) k_bi_dot2(int n, const double* a, const double* b, const double* c,
                                                          const double* d, double* part_ab, double* part_cd) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        s0 += a[i] * b[i];
        s1 += c[i] * d[i];
    }
    const double t0 = block_sum<VEC_THREADS>(s0, red);
    const double t1 = block_sum<VEC_THREADS>(s1, red);
    if (threadIdx.x == 0) { part_ab[blockIdx.x] = t0; part_cd[blockIdx.x] = t1; }
}
// x = 0, r = rhat = b, p = v = 0 ; partial of b.b
__global__ void __launch_bounds__(VEC_THREADS) k_bi_init(int n, const double* b, double* x, double* r, double* rhat, double* p,
                                                          double* v, double* part) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        const double bi = b[i];
        x[i] = 0.0; r[i] = bi; rhat[i] = bi; p[i] = 0.0; v[i] = 0.0;
        s += bi * bi;
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
// p = r + beta (p - omega v) ; y = M^-1 p
__global__ void k_bi_p(int n, double beta, double omega, const double* r, const double* v, const double* minv, double* p, double* y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double pi = r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pi;
    y[i] = pi * minv[i];
}
// s = r - alpha v (in place in r) ; z = M^-1 s
__global__ void k_bi_s(int n, double alpha, const double* v, const double* minv, double* r, double* z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double si = r[i] - alpha * v[i];
    r[i] = si;
    z[i] = si * minv[i];
}
// x += alpha y + omega z ; r = s - omega t ; partials of r.r and rhat.r
__global__ void __launch_bounds__(VEC_THREADS) k_bi_x(int n, double alpha, double omega, const double* y, const double* z,
                                                       const double* t, const double* rhat, double* x, double* r,
                                                       double* part_rr, double* part_rho) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        x[i] = x[i] + alpha * y[i] + omega * z[i];
        const double ri = r[i] - omega * t[i];
        r[i] = ri;
        s0 += ri * ri;
        s1 += rhat[i] * ri;
    }
    const double t0 = block_sum<VEC_THREADS>(s0, red);
    const double t1 = block_sum<VEC_THREADS>(s1, red);
    if (threadIdx.x == 0) { part_rr[blockIdx.x] = t0; part_rho[blockIdx.x] = t1; }
}
// iterative refinement: xt += d ; r_true = b0 - A xt is formed by the caller's SpMV into `Axt`; this kernel writes
// rnew = b0 - Axt and the partial of rnew.rnew
__global__ void __launch_bounds__(VEC_THREADS) k_bi_residual(int n, const double* b0, const double* Axt, double* rnew, double* part) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        const double ri = b0[i] - Axt[i];
        rnew[i] = ri;
        s += ri * ri;
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void k_bi_accumulate(int n, const double* d, double* xt, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xt[i] = first ? d[i] : xt[i] + d[i];
}

// the two partial arrays summed in a fixed order into out[0], out[1] (single CTA)
__global__ void __launch_bounds__(VEC_THREADS) k_bi_reduce2(const double* pa, const double* pb, int n, double* out) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    const double a = reduce_partials<VEC_THREADS>(pa, n, red);
    const double b = reduce_partials<VEC_THREADS>(pb, n, red);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

}  // namespace en234k
