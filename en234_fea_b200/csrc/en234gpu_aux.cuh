// Auxiliary kernels: batched element plugin evaluation (tests / CHECK STIFFNESS), Gauss-point state.
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

// Device twin of user_element_static / UEL for a list of elements (user_codes/src/user_element.f90:5-49):
// one CTA of 64 threads per element, thread (a,b) forms the 3x3 block K_e[a][b]; thread b < 8 also evaluates Gauss
// point b.  Output in the reference's layout: element_stiffness(ROW,COL) column-major 24x24, element_residual(24).
__global__ void __launch_bounds__(64) k_element_batch(Mesh M, int n, const int* elems /*0-based*/, const double* ut,
                                                      const double* du, double* Kout, double* Rout, CgScalars* S,
                                                      double* svout = nullptr /*n*48*/, double* enout = nullptr /*n*/) {
    constexpr int GS = GEOM_STRIDE_NEO;
    __shared__ double geom[8 * GS];
    __shared__ double sig[8 * 9];
    __shared__ double egp[8];
    const int idx = blockIdx.x;
    if (idx >= n) return;
    const int e = elems[idx];
    const int t = threadIdx.x, a = t >> 3, b = t & 7;
    const Material mat = M.mats[M.mat[e]];
    if (t < 8) {
        GpOut gp;
        double* gg = geom + t * GS;
        gauss_point_eval(M, e, t, LoadSum{ut, du}, gg, gp);
        if (!(gp.det != 0.0)) atomicOr(&S->nanflag, 1);
        double s[9];
        if (mat.kind == 1) {
            NeoPoint np;
            neo_eval(mat, gp.H, gp.det, np);
            neo_node_vectors(np, gg);
            if (!(np.J > 0.0)) atomicOr(&S->nanflag, 1);
            for (int q = 0; q < 9; ++q) s[q] = np.P[q];
        } else stress_linear(mat, gp.H, s);
        for (int q = 0; q < 9; ++q) sig[t * 9 + q] = s[q];
        // elastic strain energy density x volume at this point (abaqus_uel_linelastic_3d.for:190-191)
        const double* H = gp.H;
        egp[t] = 0.5 * (s[0] * H[0] + s[4] * H[4] + s[8] * H[8] + s[1] * (H[1] + H[3]) + s[2] * (H[2] + H[6]) +
                        s[5] * (H[5] + H[7])) * gp.det;
        if (svout) {                                              // SVARS(6*kint-5:6*kint) = stress (:193-195)
            double* sv = svout + (size_t)idx * 48 + 6 * t;
            if (mat.kind == 1) {
                NeoPoint np;
                neo_eval(mat, gp.H, gp.det, np);
                double c[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j)
                        c[3 * i + j] = (np.P[3 * i] * np.F[3 * j] + np.P[3 * i + 1] * np.F[3 * j + 1] + np.P[3 * i + 2] * np.F[3 * j + 2]) / np.J;
                sv[0] = c[0]; sv[1] = c[4]; sv[2] = c[8]; sv[3] = c[1]; sv[4] = c[2]; sv[5] = c[5];
            } else { sv[0] = s[0]; sv[1] = s[4]; sv[2] = s[8]; sv[3] = s[1]; sv[4] = s[2]; sv[5] = s[5]; }
        }
    }
    __syncthreads();
    if (t == 0 && enout) {
        double e = 0.0;
        for (int p = 0; p < 8; ++p) e += egp[p];
        enout[idx] = e;
    }
    double K[9];
    if (mat.kind == 1) {
        for (int q = 0; q < 9; ++q) K[q] = 0.0;
        for (int p = 0; p < 8; ++p) neo_block_acc(geom + p * GS, a, b, K);
    } else {
        double Sm[9];
        for (int q = 0; q < 9; ++q) Sm[q] = 0.0;
        for (int p = 0; p < 8; ++p) {
            const double* gq = geom + p * GS;
            const double w = gq[G_W];
            const double a0 = w * gq[3 * a], a1 = w * gq[3 * a + 1], a2 = w * gq[3 * a + 2];
            const double b0 = gq[3 * b], b1 = gq[3 * b + 1], b2 = gq[3 * b + 2];
            Sm[0] += a0 * b0; Sm[1] += a0 * b1; Sm[2] += a0 * b2;
            Sm[3] += a1 * b0; Sm[4] += a1 * b1; Sm[5] += a1 * b2;
            Sm[6] += a2 * b0; Sm[7] += a2 * b1; Sm[8] += a2 * b2;
        }
        const double lam = mat.d12, mu = mat.d44, tr = mu * (Sm[0] + Sm[4] + Sm[8]);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) K[3 * r + c] = lam * Sm[3 * r + c] + mu * Sm[3 * c + r] + (r == c ? tr : 0.0);
    }
    double* Ko = Kout + (size_t)idx * 576;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ko[(3 * a + r) + 24 * (3 * b + c)] = K[3 * r + c];
    if (t < 24) {
        const int na = t / 3, i = t % 3;
        double r = 0.0;
        for (int p = 0; p < 8; ++p) {
            const double* gq = geom + p * GS;
            const double* s = sig + p * 9;
            r -= (s[3 * i] * gq[3 * na] + s[3 * i + 1] * gq[3 * na + 1] + s[3 * i + 2] * gq[3 * na + 2]) * gq[G_W];
        }
        Rout[(size_t)idx * 24 + t] = r;
    }
}

// SVARS: stress s11,s22,s33,s12,s13,s23 at the 8 Gauss points (abaqus_uel_linelastic_3d.for:193-195); for the
// neo-Hookean element the Cauchy stress sigma = P F^T / J.  One thread per (element, Gauss point).
__global__ void k_svars(Mesh M, const double* ut, const double* du, double* svars) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M.E * 8) return;
    const int e = (int)(t >> 3), gp = (int)(t & 7);
    double g[GEOM_STRIDE];
    GpOut o;
    gauss_point_eval(M, e, gp, LoadSum{ut, du}, g, o);
    const Material mat = M.mats[M.mat[e]];
    double s[9];
    if (mat.kind == 1) {
        NeoPoint np;
        neo_eval(mat, o.H, o.det, np);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                s[3 * i + j] = (np.P[3 * i] * np.F[3 * j] + np.P[3 * i + 1] * np.F[3 * j + 1] + np.P[3 * i + 2] * np.F[3 * j + 2]) / np.J;
    } else stress_linear(mat, o.H, s);
    double* sv = svars + (size_t)e * 48 + 6 * gp;
    sv[0] = s[0]; sv[1] = s[4]; sv[2] = s[8]; sv[3] = s[1]; sv[4] = s[2]; sv[5] = s[5];
}

// FP64 FMA throughput of the device (the roof of the element kernel and of the general-mesh matrix-free operator):
// 16 independent FMA chains per thread, `iters` x 16 FMAs each, full occupancy.  flops = 2 * 16 * iters * threads.
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double a, double* out) {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = fma(v[k], a, 1e-9);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    if (s == 12345.678) out[0] = s;                              // keeps the chains alive
}

}  // namespace en234k
