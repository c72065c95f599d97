// Matrix-free element-by-element operator  y = sum_e P_e^T K_e P_e x  (BASELINE config 5; the reference has no such
// path — its CG SpMV, src/solver_conjugate_gradient.f90:877-883, always reads an assembled matrix).
// K_e x_e is evaluated as the internal force of the linear-elastic element for the displacement field x
// (the integrand of el_linelast_3Dbasic.f90:124-128 with u := x), never forming K_e.
// Determinism without atomics: elements are greedily coloured so that no two elements of a colour share a node;
// colours are processed by successive launches, each element of a colour read-modify-writes its 8 node rows.
#pragma once
#include <algorithm>
#include <vector>

#include "en234gpu_kernels.cuh"

namespace en234k {

struct MfData {
    int n_colors = 0;
    std::vector<int> color_ptr;     // host
    int* d_elist = nullptr;         // elements grouped by colour, ascending within a colour
    int n_partials = 0;
    int launches_per_apply = 0;
    int grid_vec = 0;
};

// 8 lanes = the 8 Gauss points of one element.  mode 0: y += K_e x (x masked by cflag);  mode 1: diag(K_e) into y.
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

// y_row = x_row on prescribed rows (unit diagonal on the owner rank only), partial of x.y over all local rows
__global__ void __launch_bounds__(VEC_THREADS) k_mf_finish(int n, const uint8_t* cflag, const uint8_t* owned, const double* x,
                                                            double* y, double* partials, const CgScalars* S) {
    __shared__ double red[VEC_THREADS / 32 + 1];
    if (S && S->stop) return;
    double dot = 0.0;
    for (int d = blockIdx.x * VEC_THREADS + threadIdx.x; d < n; d += gridDim.x * VEC_THREADS) {
        double v = y[d];
        if (cflag && cflag[d]) { v = (owned == nullptr || owned[d / 3]) ? x[d] : 0.0; y[d] = v; }
        dot += v * x[d];
    }
    const double t = block_sum<VEC_THREADS>(dot, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void k_mask_values(int n, const uint8_t* cflag, const double* cval, double* out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) out[d] = cflag[d] ? cval[d] : 0.0;
}
__global__ void k_fix_diag(int n, const uint8_t* cflag, const uint8_t* owned, double* diag) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n && cflag[d]) diag[d] = (owned == nullptr || owned[d / 3]) ? 1.0 : 0.0;
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
    mf.n_partials = mf.grid_vec;
    mf.launches_per_apply = ncol + 1;
    (void)M;
    return 0;
}
inline void mf_release(MfData& mf) { if (mf.d_elist) cudaFree(mf.d_elist); mf.d_elist = nullptr; }

// y = P A P x + (I-P) x ; partials of x.y.  (cflag == null: plain A x)
inline void mf_apply(const MfData& mf, const Mesh& M, int ndof, const uint8_t* cflag, const uint8_t* owned, const double* x,
                     double* y, double* partials, const CgScalars* S, cudaStream_t s, long long* launches) {
    cudaMemsetAsync(y, 0, (size_t)ndof * sizeof(double), s);
    for (int c = 0; c < mf.n_colors; ++c) {
        const int n = mf.color_ptr[c + 1] - mf.color_ptr[c];
        if (n == 0) continue;
        k_mf_color<0><<<(8 * n + 255) / 256, 256, 0, s>>>(M, n, mf.d_elist + mf.color_ptr[c], cflag, x, y, S);
    }
    k_mf_finish<<<mf.grid_vec, VEC_THREADS, 0, s>>>(ndof, cflag, owned, x, y, partials, S);
    *launches += mf.n_colors + 1;
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
