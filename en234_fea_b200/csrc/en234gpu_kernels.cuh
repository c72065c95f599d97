// Device kernels of the EN234_FEA static-step hot path for B200 (sm_100a), FP64.
// Reference arithmetic being restated (paths relative to /root/reference/EN234_FEA/):
//   element:  user_codes/src/el_linelast_3Dbasic.f90:79-133 = user_codes/src/abaqus_uel_linelastic_3d.for:129-196
//   Gauss:    user_codes/modules/Element_Utilities.f90:636-649 (single-precision abscissa literal)
//   shape:    Element_Utilities.f90:908-940, inverse :99-123
//   assembly: src/solver_conjugate_gradient.f90:233-258 (summed in ascending element order)
//   BCs:      src/solver_conjugate_gradient.f90:673-713 (fixdof), :417-671
//   PCG:      src/solver_conjugate_gradient.f90:776-912
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace en234k {

constexpr int WARPS = 8;               // warps per CTA in the row kernels
constexpr int THREADS = WARPS * 32;
constexpr int MAXNB = 32;              // block-row length handled in shared memory by the assembly kernel
constexpr int GEOM_STRIDE = 25;        // per Gauss point: dN/dx (8x3) + w|J|
constexpr int GEOM_PER_WARP = 4 * 8 * GEOM_STRIDE;
constexpr int ACC_PER_WARP = MAXNB * 9;
constexpr int VEC_THREADS = 256;
constexpr int MAX_PARTIALS = 4096;

// (double)0.5773502692f : the value the reference actually integrates with
#define EN234_GP 0.57735025882720947

struct Material {                      // one entry per distinct (element flag, property set)
    int kind;                          // 0 linear elastic (1001 / U1), 1 neo-Hookean (1002 / U2)
    double d11, d12, d44;              // el_linelast_3Dbasic.f90:96-98
    double mu, kappa;                  // neo-Hookean
};

struct CgScalars {                     // device-resident PCG state (no host round trip inside an iteration)
    double rz[2];                      // z.r of the current / next iteration (ping-pong on iteration parity)
    double bnrm;                       // b.b
    double err;                        // (r.r)/(b.b)
    double pAp;
    double tol;
    double corr2;                      // sol.sol
    double norm2;                      // scratch for norms
    int iter, maxit, stop, fail;
    int nanflag, pad0, pad1, pad2;
};

struct Mesh {
    int N, E;
    const double *x, *y, *z;           // SoA node coordinates
    const int* conn;                   // SoA connectivity conn[a*E + e], 0-based nodes
    const int* mat;                    // material index per element
    const Material* mats;
    int n_mats;                        // 1: every element uses mats[0] (kernels then skip the per-element lookup)
    const int* adj_ptr;                // node -> (element, local node) adjacency, ascending element
    const int* adj;                    // e*8 + a
    const int* adj_pos;                // inverse of adj: adjacency position of (element e, local node a) at [8e + a]
    const uint8_t* slot;               // 8 per adjacency entry: position of column conn[e][b] in the node's block row
    const int* browptr;                // block-row pointers (N+1)
    const int* bcol;                   // block columns, ascending within a row
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still draining; pdl_wait() blocks until the predecessor grid has completed and its
// writes are visible (a no-op for ordinary launches), pdl_trigger() lets the successor start launching.  Everything a
// kernel reads before pdl_wait() must not be written by its immediate predecessor.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum of three values over the warp in 6 shuffle steps instead of 15: the first two steps halve what a lane carries
// (xor 16: keep (y0,y1) or y2; xor 8: keep one of the pair), the last three are an ordinary butterfly.  The sum of y_r ends
// up in lanes 8r .. 8r+7 (fixed tree: deterministic).
__device__ __forceinline__ double warp_sum3(double y0, double y1, double y2, int lane) {
    const bool hi16 = lane & 16, hi8 = lane & 8;
    const double sa = hi16 ? y0 : y2, sb = hi16 ? y1 : 0.0;          // what the partner keeps
    double ka = (hi16 ? y2 : y0) + __shfl_xor_sync(0xffffffffu, sa, 16);
    double kb = (hi16 ? 0.0 : y1) + __shfl_xor_sync(0xffffffffu, sb, 16);
    double v = (hi8 ? kb : ka) + __shfl_xor_sync(0xffffffffu, hi8 ? ka : kb, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
// deterministic CTA sum (fixed tree); result valid in every thread
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sm /*NT/32 + 1*/) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += sm[w];
        sm[NT / 32] = s;
    }
    __syncthreads();
    return sm[NT / 32];
}
// every CTA reduces the same partial array in the same order -> identical scalar everywhere
template <int NT>
__device__ __forceinline__ double reduce_partials(const double* __restrict__ part, int n, double* sm) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += NT) s += part[i];
    return block_sum<NT>(s, sm);
}

// sign of node c in direction j (ABAQUS C3D8 order, Element_Utilities.f90:908-916) as a compile-time table
__device__ __forceinline__ constexpr int sgn(int c, int j) {
    return j == 0 ? ((c == 1 || c == 2 || c == 5 || c == 6) ? 1 : -1)
         : j == 1 ? ((c == 2 || c == 3 || c == 6 || c == 7) ? 1 : -1)
                  : ((c >= 4) ? 1 : -1);
}

// dN_c/dxi at the Gauss point whose xi_j sign bits are (gp>>j)&1: (1 +- xi)(1 +- xi)/8 with the sign applied last,
// the same operation order as Element_Utilities.f90:916-939.
template <int C>
__device__ __forceinline__ void dshape(const double fp[3], const double fm[3], double dn[3]) {
    // fp[j] = 1 + xi_j, fm[j] = 1 - xi_j
    const double f0 = sgn(C, 0) > 0 ? fp[0] : fm[0];
    const double f1 = sgn(C, 1) > 0 ? fp[1] : fm[1];
    const double f2 = sgn(C, 2) > 0 ? fp[2] : fm[2];
    dn[0] = sgn(C, 0) * (f1 * f2 * 0.125);
    dn[1] = sgn(C, 1) * (f0 * f2 * 0.125);
    dn[2] = sgn(C, 2) * (f0 * f1 * 0.125);
}

struct Jac { double Ji[9]; double det; };

// cofactor inverse, Element_Utilities.f90:99-123 ; J[3*i+j] = dx_i/dxi_j
__device__ __forceinline__ void invert3(const double* J, Jac& o) {
    const double det = J[0] * J[4] * J[8] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8] + J[1] * J[5] * J[6] +
                       J[2] * J[3] * J[7] - J[2] * J[4] * J[6];
    o.det = det;
    const double r = 1.0 / det;
    o.Ji[0] = (J[4] * J[8] - J[5] * J[7]) * r;
    o.Ji[1] = -(J[1] * J[8] - J[2] * J[7]) * r;
    o.Ji[2] = (J[1] * J[5] - J[2] * J[4]) * r;
    o.Ji[3] = -(J[3] * J[8] - J[5] * J[6]) * r;
    o.Ji[4] = (J[0] * J[8] - J[2] * J[6]) * r;
    o.Ji[5] = -(J[0] * J[5] - J[2] * J[3]) * r;
    o.Ji[6] = (J[3] * J[7] - J[4] * J[6]) * r;
    o.Ji[7] = -(J[0] * J[7] - J[1] * J[6]) * r;
    o.Ji[8] = (J[0] * J[4] - J[1] * J[3]) * r;
}

// Geometry + displacement gradient of element e at Gauss point gp (one lane = one Gauss point).
// Writes dN_c/dx (c = 0..7) and w|J| to g[0..24]; returns grad u (H[3*i+j] = du_i/dx_j) and det.
struct GpOut { double H[9]; double det; };

template <int C>
__device__ __forceinline__ void jac_acc(const Mesh& M, int e, const double fp[3], const double fm[3], double* J) {
    const int n = M.conn[(size_t)C * M.E + e];
    const double xc = __ldg(M.x + n), yc = __ldg(M.y + n), zc = __ldg(M.z + n);
    double dn[3];
    dshape<C>(fp, fm, dn);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        J[j] += xc * dn[j];
        J[3 + j] += yc * dn[j];
        J[6 + j] += zc * dn[j];
    }
}
// how the nodal vector whose gradient is taken is read
struct LoadSum {                       // u = dof_total + dof_increment                 (solver_direct.f90:234)
    const double* __restrict__ ut;
    const double* __restrict__ du;
    __device__ __forceinline__ double operator()(size_t d) const { return ut[d] + du[d]; }
    __device__ __forceinline__ bool stopped() const { return false; }
    bool may_skip() const { return false; }
    static constexpr bool use_geometry_cache = false;
};
struct LoadMasked {                    // x with prescribed DOFs zeroed (operator P A P of the matrix-free solve)
    const double* __restrict__ x;
    const uint8_t* __restrict__ cflag; // may be null
    const CgScalars* stop = nullptr;   // PCG stop flag: kernels of a captured iteration exit at once when set
    // both loads are issued unconditionally: a load of x behind the branch on cflag would cost two memory latencies
    __device__ __forceinline__ double operator()(size_t d) const {
        const double v = x[d];
        const uint8_t f = cflag ? cflag[d] : (uint8_t)0;
        return f ? 0.0 : v;
    }
    __device__ __forceinline__ bool stopped() const { return stop != nullptr && stop->stop != 0; }
    bool may_skip() const { return stop != nullptr; }
    static constexpr bool use_geometry_cache = true;
};

template <int C, class Load>
__device__ __forceinline__ void grad_acc(const Mesh& M, int e, const double fp[3], const double fm[3], const double* Ji,
                                         const Load& ld, double* g, double* H) {
    double dn[3], d[3];
    dshape<C>(fp, fm, dn);
#pragma unroll
    for (int j = 0; j < 3; ++j) d[j] = dn[0] * Ji[j] + dn[1] * Ji[3 + j] + dn[2] * Ji[6 + j];
    g[3 * C] = d[0]; g[3 * C + 1] = d[1]; g[3 * C + 2] = d[2];
    const int n = M.conn[(size_t)C * M.E + e];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double u = ld(3 * (size_t)n + i);
#pragma unroll
        for (int j = 0; j < 3; ++j) H[3 * i + j] += u * d[j];
    }
}

template <class Load>
__device__ __forceinline__ void gauss_point_eval(const Mesh& M, int e, int gp, const Load& ld, double* g, GpOut& o) {
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
    jac_acc<0>(M, e, fp, fm, J); jac_acc<1>(M, e, fp, fm, J); jac_acc<2>(M, e, fp, fm, J); jac_acc<3>(M, e, fp, fm, J);
    jac_acc<4>(M, e, fp, fm, J); jac_acc<5>(M, e, fp, fm, J); jac_acc<6>(M, e, fp, fm, J); jac_acc<7>(M, e, fp, fm, J);
    Jac jc;
    invert3(J, jc);
    o.det = jc.det;
#pragma unroll
    for (int q = 0; q < 9; ++q) o.H[q] = 0.0;
    grad_acc<0>(M, e, fp, fm, jc.Ji, ld, g, o.H); grad_acc<1>(M, e, fp, fm, jc.Ji, ld, g, o.H);
    grad_acc<2>(M, e, fp, fm, jc.Ji, ld, g, o.H); grad_acc<3>(M, e, fp, fm, jc.Ji, ld, g, o.H);
    grad_acc<4>(M, e, fp, fm, jc.Ji, ld, g, o.H); grad_acc<5>(M, e, fp, fm, jc.Ji, ld, g, o.H);
    grad_acc<6>(M, e, fp, fm, jc.Ji, ld, g, o.H); grad_acc<7>(M, e, fp, fm, jc.Ji, ld, g, o.H);
    g[24] = jc.det;   // w = 1 for the 2x2x2 rule
}

// Cauchy/PK1-type stress measure at a Gauss point.
//  linear elastic: sigma (symmetric 3x3 in s[0..8]) = D * strain(H)            (el_linelast_3Dbasic.f90:124-127)
//  neo-Hookean:    first Piola-Kirchhoff P(F = I + H) and tangent scalars        (oracle/or_elements.cpp el_neohooke_3d)
// Per-Gauss-point layout in shared memory (doubles): dN/dX of the 8 nodes, w|J0|, then (neo-Hookean only)
// alpha_c = F^-T dN_c, f_c = F dN_c and four tangent scalars.
constexpr int G_DN = 0, G_W = 24, G_AL = 25, G_FF = 49, G_SC = 73;
constexpr int GEOM_STRIDE_NEO = 77;

// Compressible neo-Hookean point quantities (oracle/or_elements.cpp el_neohooke_3d; W = mu/2 (Ibar1-3) + K/2 (J-1)^2):
//   P_iJ = mu J^-2/3 (F_iJ - I1/3 F^-1_Ji) + K (J-1) J F^-1_Ji
//   A_iJkL contracted with dN_a,J dN_b,L:
//   K_ik = s0 d_ik (a.b) - s1 (fa_i beta_k + alpha_i fb_k) + s2 alpha_i beta_k + s3 alpha_k beta_i
//   s0 = w c1, s1 = w 2/3 c1, s2 = w (2/9 I1 c1 + K(2J-1)J), s3 = w (c1 I1/3 - K(J-1)J),  c1 = mu J^-2/3
struct NeoPoint { double F[9], Fi[9], P[9], sc[4], J; };

__device__ __forceinline__ void neo_eval(const Material& m, const double* H, double w, NeoPoint& o) {
#pragma unroll
    for (int q = 0; q < 9; ++q) o.F[q] = H[q] + ((q == 0 || q == 4 || q == 8) ? 1.0 : 0.0);
    Jac jc;
    invert3(o.F, jc);
    o.J = jc.det;
    double I1 = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) { o.Fi[q] = jc.Ji[q]; I1 += o.F[q] * o.F[q]; }
    const double t = rcbrt(o.J);
    const double c1 = m.mu * t * t;
    const double cJ = m.kappa * (2.0 * o.J - 1.0) * o.J, cK = m.kappa * (o.J - 1.0) * o.J;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.P[3 * i + j] = c1 * (o.F[3 * i + j] - I1 / 3.0 * o.Fi[3 * j + i]) + cK * o.Fi[3 * j + i];
    o.sc[0] = w * c1;
    o.sc[1] = w * (2.0 / 3.0) * c1;
    o.sc[2] = w * ((2.0 / 9.0) * I1 * c1 + cJ);
    o.sc[3] = w * (c1 * I1 / 3.0 - cK);
}
// alpha_c, f_c for the 8 nodes from dN/dX stored at g[0..23]
__device__ __forceinline__ void neo_node_vectors(const NeoPoint& p, double* g) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double d0 = g[3 * c], d1 = g[3 * c + 1], d2 = g[3 * c + 2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g[G_AL + 3 * c + i] = p.Fi[i] * d0 + p.Fi[3 + i] * d1 + p.Fi[6 + i] * d2;
            g[G_FF + 3 * c + i] = p.F[3 * i] * d0 + p.F[3 * i + 1] * d1 + p.F[3 * i + 2] * d2;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) g[G_SC + q] = p.sc[q];
}
// one Gauss point's contribution to the neo-Hookean block K[a][b]
__device__ __forceinline__ void neo_block_acc(const double* gq, int a, int b, double* K) {
    const double a0 = gq[3 * a], a1 = gq[3 * a + 1], a2 = gq[3 * a + 2];
    const double b0 = gq[3 * b], b1 = gq[3 * b + 1], b2 = gq[3 * b + 2];
    const double al[3] = {gq[G_AL + 3 * a], gq[G_AL + 3 * a + 1], gq[G_AL + 3 * a + 2]};
    const double be[3] = {gq[G_AL + 3 * b], gq[G_AL + 3 * b + 1], gq[G_AL + 3 * b + 2]};
    const double fa[3] = {gq[G_FF + 3 * a], gq[G_FF + 3 * a + 1], gq[G_FF + 3 * a + 2]};
    const double fb[3] = {gq[G_FF + 3 * b], gq[G_FF + 3 * b + 1], gq[G_FF + 3 * b + 2]};
    const double s0 = gq[G_SC], s1 = gq[G_SC + 1], s2 = gq[G_SC + 2], s3 = gq[G_SC + 3];
    const double ab = s0 * (a0 * b0 + a1 * b1 + a2 * b2);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            K[3 * i + k] += s2 * al[i] * be[k] + s3 * al[k] * be[i] - s1 * (fa[i] * be[k] + al[i] * fb[k]) + (i == k ? ab : 0.0);
}

__device__ __forceinline__ void stress_linear(const Material& m, const double* H, double* s) {
    const double e11 = H[0], e22 = H[4], e33 = H[8];
    const double g12 = H[1] + H[3], g13 = H[2] + H[6], g23 = H[5] + H[7];
    s[0] = m.d11 * e11 + m.d12 * e22 + m.d12 * e33;
    s[4] = m.d12 * e11 + m.d11 * e22 + m.d12 * e33;
    s[8] = m.d12 * e11 + m.d12 * e22 + m.d11 * e33;
    s[1] = s[3] = m.d44 * g12;
    s[2] = s[6] = m.d44 * g13;
    s[5] = s[7] = m.d44 * g23;
}

}  // namespace en234k
