// Jacobi-preconditioned BiCGSTAB for UNSYMMETRIC stiffness matrices (SOLVER, ..., UNSYMMETRIC: src/solver_direct.f90
// stores alow for these, :282, :952-958; the reference's CG solver cannot handle them).  Used by the built-in C3D8
// continuum elements.  Like the PCG loop, the recurrence scalars live in device memory
// (BiScalars), every kernel of an iteration starts by testing the stop flag, and a chunk of iterations is one CUDA graph:
// no host round trip inside an iteration.  Deterministic: per-CTA partials reduced in a fixed order by one CTA.
//   rho = rhat.r ; beta = (rho/rho_old)(alpha/omega) ; p = r + beta (p - omega v) ; y = M^-1 p ; v = A y
//   alpha = rho / rhat.v ; s = r - alpha v ; z = M^-1 s ; t = A z ; omega = t.s / t.t
//   x += alpha y + omega z ; r = s - omega t ; stop when (r.r)/(b.b) <= tol
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

struct BiScalars {
    double rho, rho_old, alpha, omega, rr, bb, tol;
    int iter, maxit, stop, fail;
};

// x = 0, r = rhat = b, p = v = 0 ; partial of b.b
// (owned: per-NODE ownership mask of partitioned runs — every dot product counts each global DOF once; null = all)
__device__ __forceinline__ bool bi_owned(const uint8_t* owned, int i) { return owned == nullptr || owned[i / 3] != 0; }
__global__ void __launch_bounds__(VEC_THREADS) k_bi_init(int n, const double* b, double* x, double* r, double* rhat, double* p,
                                                          double* v, double* part, const uint8_t* owned = nullptr) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        const double bi = b[i];
        x[i] = 0.0; r[i] = bi; rhat[i] = bi; p[i] = 0.0; v[i] = 0.0;
        if (bi_owned(owned, i)) s += bi * bi;
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(VEC_THREADS) k_bi_init_finish(const double* part, int n, BiScalars* S, double tol, int maxit) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    const double bb = reduce_partials<VEC_THREADS>(part, n, red);
    if (threadIdx.x == 0) {
        S->bb = bb; S->rr = bb; S->rho = bb; S->rho_old = 1.0; S->alpha = 1.0; S->omega = 1.0; S->tol = tol;
        S->iter = 0; S->maxit = maxit;
        S->fail = (bb == bb) ? 0 : 1;
        S->stop = (bb > 0.0) ? 0 : 1;                 // zero right-hand side: x = 0
    }
}
// p = r + beta (p - omega v) ; y = M^-1 p
__global__ void k_bi_p(int n, const BiScalars* S, const double* r, const double* v, const double* minv, double* p, double* y) {
    if (S->stop) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double beta = S->iter == 0 ? 0.0 : (S->rho / S->rho_old) * (S->alpha / S->omega);
    const double pi = r[i] + beta * (p[i] - S->omega * v[i]);
    p[i] = pi;
    y[i] = pi * minv[i];
}
// partials of a.b and c.d
__global__ void __launch_bounds__(VEC_THREADS) k_bi_dot2(int n, const BiScalars* S, const double* a, const double* b,
                                                          const double* c, const double* d, double* part_ab, double* part_cd,
                                                          const uint8_t* owned = nullptr) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S && S->stop) return;
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        if (bi_owned(owned, i)) {
            s0 += a[i] * b[i];
            s1 += c[i] * d[i];
        }
    }
    const double t0 = block_sum<VEC_THREADS>(s0, red);
    const double t1 = block_sum<VEC_THREADS>(s1, red);
    if (threadIdx.x == 0) { part_ab[blockIdx.x] = t0; part_cd[blockIdx.x] = t1; }
}
// alpha = rho / (rhat.v)
__global__ void __launch_bounds__(VEC_THREADS) k_bi_alpha(const double* part, int n, BiScalars* S) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S->stop) return;
    const double rv = reduce_partials<VEC_THREADS>(part, n, red);
    if (threadIdx.x == 0) {
        if (rv == 0.0 || rv != rv) { S->fail = 1; S->stop = 1; }
        else S->alpha = S->rho / rv;
    }
}
// s = r - alpha v (in place in r) ; z = M^-1 s
__global__ void k_bi_s(int n, const BiScalars* S, const double* v, const double* minv, double* r, double* z) {
    if (S->stop) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double si = r[i] - S->alpha * v[i];
    r[i] = si;
    z[i] = si * minv[i];
}
// omega = (t.s) / (t.t)
__global__ void __launch_bounds__(VEC_THREADS) k_bi_omega(const double* part_ts, const double* part_tt, int n, BiScalars* S) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S->stop) return;
    const double ts = reduce_partials<VEC_THREADS>(part_ts, n, red);
    const double tt = reduce_partials<VEC_THREADS>(part_tt, n, red);
    if (threadIdx.x == 0) {
        if (tt != tt) { S->fail = 1; S->stop = 1; }
        else S->omega = tt > 0.0 ? ts / tt : 0.0;
    }
}
// x += alpha y + omega z ; r = s - omega t ; partials of r.r and rhat.r
__global__ void __launch_bounds__(VEC_THREADS) k_bi_x(int n, const BiScalars* S, const double* y, const double* z,
                                                       const double* t, const double* rhat, double* x, double* r,
                                                       double* part_rr, double* part_rho, const uint8_t* owned = nullptr) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S->stop) return;
    const double alpha = S->alpha, omega = S->omega;
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        x[i] = x[i] + alpha * y[i] + omega * z[i];
        const double ri = r[i] - omega * t[i];
        r[i] = ri;
        if (bi_owned(owned, i)) {
            s0 += ri * ri;
            s1 += rhat[i] * ri;
        }
    }
    const double t0 = block_sum<VEC_THREADS>(s0, red);
    const double t1 = block_sum<VEC_THREADS>(s1, red);
    if (threadIdx.x == 0) { part_rr[blockIdx.x] = t0; part_rho[blockIdx.x] = t1; }
}
// end of an iteration: rho_old = rho, rho = rhat.r, iteration count, stopping rule, breakdown / iteration cap
__global__ void __launch_bounds__(VEC_THREADS) k_bi_finish(const double* part_rr, const double* part_rho, int n, BiScalars* S) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S->stop) return;
    const double rr = reduce_partials<VEC_THREADS>(part_rr, n, red);
    const double rho = reduce_partials<VEC_THREADS>(part_rho, n, red);
    if (threadIdx.x == 0) {
        S->rho_old = S->rho;
        S->rho = rho;
        S->rr = rr;
        S->iter = S->iter + 1;
        if (rr != rr) { S->fail = 1; S->stop = 1; }
        else if (!(rr / S->bb > S->tol)) S->stop = 1;
        else if (S->iter >= S->maxit || S->omega == 0.0 || rho == 0.0) { S->fail = 1; S->stop = 1; }
    }
}

// iterative refinement: xt += d ; r_true = b0 - A xt is formed by the caller's SpMV into `Axt`; this kernel writes
// rnew = b0 - Axt and the partials of rnew.rnew and b0.b0
__global__ void __launch_bounds__(VEC_THREADS) k_bi_residual(int n, const double* b0, const double* Axt, double* rnew,
                                                              double* part_rr, double* part_bb, const uint8_t* owned = nullptr) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    double s = 0.0, sb = 0.0;
    for (int i = blockIdx.x * VEC_THREADS + threadIdx.x; i < n; i += gridDim.x * VEC_THREADS) {
        const double bi = b0[i], ri = bi - Axt[i];
        rnew[i] = ri;
        if (bi_owned(owned, i)) {
            s += ri * ri;
            sb += bi * bi;
        }
    }
    const double t = block_sum<VEC_THREADS>(s, red);
    const double tb = block_sum<VEC_THREADS>(sb, red);
    if (threadIdx.x == 0) { part_rr[blockIdx.x] = t; part_bb[blockIdx.x] = tb; }
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
