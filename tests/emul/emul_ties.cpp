// TEST INFRASTRUCTURE: compiles the tie-constraint device kernels (en234_fea_b200/csrc/en234gpu_ties.cuh) for the HOST, by
// defining the CUDA keywords away, and runs the device path's elimination — known member corrections to the right-hand side,
// reduction of right-hand side and Jacobi diagonal, Jacobi-PCG with the expand / reduce wrappers around a DENSE matrix-vector
// product, member corrections, the tied rows of rhs - K du — thread by thread under the address / UB sanitisers.
// tests/test_oracle_cpu.py compares the result with a dense solve of the saddle-point system.  Not part of the product.
//   usage: emul_ties <in.bin> <out.bin>
//   in : int32 n, nt, ng, nm, ntd; f64 K[n*n] (row-major, boundary conditions applied); f64 rhs[n]; f64 minv[n]; f64 ut[n]; f64 du[n];
//        f64 cval[n]; uint8 as int32 cflag[n]; int32 d1[nt], d2[nt]; f64 lambda[nt];
//        int32 g_root[ng], g_ptr[ng+1], g_fixed[ng]; int32 m_dof[nm], m_root[nm], m_fixed[nm]; int32 tdof[ntd], tptr[ntd+1], tent[2*nt]
//   out: f64 sol[n]; f64 trow[ntd]; f64 gap[nt]; f64 rhs_tied[n] (k_tie_rhs applied to rhs); f64 iterations
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#define __CUDA_RUNTIME_H__
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __ldg(p) (*(p))
#define __ldcs(p) (*(p))
#define __shfl_xor_sync(m, v, o) (v)
#define __shfl_down_sync(m, v, o) (0.0)
#define __syncthreads()
#define __syncwarp()
#define __threadfence()
static inline double rcbrt(double x) { return 1.0 / std::cbrt(x); }
struct D3 { int x = 0, y = 0, z = 0; };
static D3 blockIdx, threadIdx, gridDim, blockDim;
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
#include "en234gpu_ties.cuh"
using namespace en234k;

template <class T> static void rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
}
// run a thread-per-item kernel over `items` items with 256-thread blocks, one thread at a time
template <class F> static void launch(int items, F&& body) {
    blockDim.x = 256;
    gridDim.x = std::max(1, (items + 255) / 256);
    for (int b = 0; b < gridDim.x; ++b)
        for (int t = 0; t < 256; ++t) { blockIdx.x = b; threadIdx.x = t; body(); }
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<int> hdr, cflag_i, d1, d2, g_root, g_ptr, g_fixed_i, m_dof, m_root, m_fixed_i, tdof, tptr, tent;
    std::vector<double> K, rhs, minv, ut, du, cval, lambda;
    rd(f, hdr, 5);
    const int n = hdr[0], nt = hdr[1], ng = hdr[2], nm = hdr[3], ntd = hdr[4];
    rd(f, K, (size_t)n * n); rd(f, rhs, n); rd(f, minv, n); rd(f, ut, n); rd(f, du, n); rd(f, cval, n); rd(f, cflag_i, n);
    rd(f, d1, nt); rd(f, d2, nt); rd(f, lambda, nt);
    rd(f, g_root, ng); rd(f, g_ptr, ng + 1); rd(f, g_fixed_i, ng); rd(f, m_dof, nm); rd(f, m_root, nm); rd(f, m_fixed_i, nm);
    rd(f, tdof, ntd); rd(f, tptr, ntd + 1); rd(f, tent, 2 * (size_t)nt);
    std::fclose(f);
    std::vector<uint8_t> cflag(cflag_i.begin(), cflag_i.end()), g_fixed(g_fixed_i.begin(), g_fixed_i.end()), m_fixed(m_fixed_i.begin(), m_fixed_i.end());
    const int counts[2] = {ng, nm};
    TieTables T;
    T.counts = counts; T.g_root = g_root.data(); T.g_ptr = g_ptr.data(); T.g_fixed = g_fixed.data();
    T.m_dof = m_dof.data(); T.m_root = m_root.data(); T.m_fixed = m_fixed.data();
    auto matvec = [&](const std::vector<double>& x, std::vector<double>& y) {
        for (int i = 0; i < n; ++i) { double s = 0.0; for (int j = 0; j < n; ++j) s += K[(size_t)i * n + j] * x[j]; y[i] = s; }
    };
    // ---- k_tie_rhs / k_tie_gap on copies (the assembly-side kernels)
    std::vector<double> rhs_tied = rhs, rforce(n, 0.0), gap(nt, 0.0);
    launch(ntd, [&] { k_tie_rhs(ntd, tdof.data(), tptr.data(), tent.data(), lambda.data(), rhs_tied.data(), rforce.data()); });
    for (int i = 0; i < ntd; ++i) if (rforce[tdof[i]] != -rhs_tied[tdof[i]]) return 3;
    launch(nt, [&] { k_tie_gap(nt, d1.data(), d2.data(), ut.data(), du.data(), cflag.data(), cval.data(), gap.data()); });
    // ---- solve: the sequence of ties_before_solve / PCG with wrappers / ties_after_solve (en234gpu_ties_host.h)
    std::vector<double> rhs0 = rhs, d0(n, 0.0), mgap(std::max(nm, 1), 0.0), vec(n, 0.0);
    launch(ntd, [&] { k_tie_d0(T, ut.data(), du.data(), cval.data(), d0.data(), mgap.data()); });
    matvec(d0, vec);
    blockDim.x = 256; gridDim.x = 4;
    for (int b = 0; b < 4; ++b) for (int t = 0; t < 256; ++t) { blockIdx.x = b; threadIdx.x = t; k_vec_sub(n, vec.data(), rhs.data()); }
    launch(ntd, [&] { k_tie_reduce_rhs(T, rhs.data(), minv.data()); });
    std::vector<double> sol(n, 0.0), r = rhs, z(n), p(n), Ap(n);
    double bb = 0.0, rz = 0.0;
    for (int i = 0; i < n; ++i) { z[i] = minv[i] * r[i]; p[i] = z[i]; bb += rhs[i] * rhs[i]; rz += r[i] * z[i]; }
    int it = 0;
    for (; it < 20 * n && bb > 0.0; ++it) {
        launch(ntd, [&] { k_tie_expand(T, p.data(), nullptr); });
        matvec(p, Ap);
        double pAp = 0.0;
        for (int i = 0; i < n; ++i) pAp += p[i] * Ap[i];        // like k_spmv: before the reduction, expanded p
        launch(ntd, [&] { k_tie_reduce(T, Ap.data(), p.data(), nullptr); });
        const double alpha = rz / pAp;
        double rr = 0.0, rz_new = 0.0;
        for (int i = 0; i < n; ++i) { sol[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rr += r[i] * r[i]; z[i] = minv[i] * r[i]; rz_new += r[i] * z[i]; }
        if (rr / bb < 1e-30) { ++it; break; }
        const double beta = rz_new / rz;
        rz = rz_new;
        for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    launch(ntd, [&] { k_tie_expand_sol(T, mgap.data(), sol.data()); });
    matvec(sol, vec);
    std::vector<double> trow(ntd, 0.0);
    launch(ntd, [&] { k_tie_rows(ntd, tdof.data(), rhs0.data(), vec.data(), trow.data()); });
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    const double iters = it;
    std::fwrite(sol.data(), 8, n, o); std::fwrite(trow.data(), 8, ntd, o); std::fwrite(gap.data(), 8, nt, o);
    std::fwrite(rhs_tied.data(), 8, n, o); std::fwrite(&iters, 8, 1, o);
    std::fclose(o);
    return 0;
}
