// TEST INFRASTRUCTURE: compiles the thread-per-node device kernels of en234_fea_b200/csrc (state-print projection, explicit
// dynamics) for the HOST, by defining the CUDA keywords away, and runs them thread by thread on a mesh read from a file.
// tests/test_oracle_cpu.py builds it with g++ -fsanitize=address and compares the outputs with the oracle: the kernels'
// arithmetic and indexing are then checked in the CPU suite (bounds errors included), on the same source nvcc compiles.
// Not part of the product: nothing under en234_fea_b200/ refers to it.
//   usage: emul_kernels <in.bin> <out.bin>
//   in : int32 N, E, nf, source, nsp; int32 codes[nf]; f64 coords[3N]; int32 conn[8E] (1-based); f64 props[3] (d11,d12,d44);
//        f64 ut[3N], du[3N]; f64 density[E]; f64 sv[E*8*nsp] (source 1)
//   out: f64 F[nf*N]; f64 mdiag[N]; f64 lump[N]; f64 rowsum[N] (sum of the node's matrix row); f64 mass[3N]
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
#define __ldg(p) (*(p))
#define __ldcs(p) (*(p))
#define __shfl_xor_sync(m, v, o) (v)
#define __syncthreads()
#define __syncwarp()
#define __threadfence()
static inline double rcbrt(double x) { return 1.0 / std::cbrt(x); }
struct D3 { int x = 0, y = 0, z = 0; };
static D3 blockIdx, threadIdx, gridDim, blockDim;
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
#include "en234gpu_project.cuh"
#include "en234gpu_dyn.cuh"
using namespace en234k;

template <class T> static void rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<int> hdr, codes, connectivity;
    std::vector<double> coords, props, ut, du, density, sv;
    rd(f, hdr, 5);
    const int N = hdr[0], E = hdr[1], nf = hdr[2], source = hdr[3], nsp = hdr[4];
    rd(f, codes, nf); rd(f, coords, 3 * (size_t)N); rd(f, connectivity, 8 * (size_t)E); rd(f, props, 3);
    rd(f, ut, 3 * (size_t)N); rd(f, du, 3 * (size_t)N); rd(f, density, E);
    rd(f, sv, source == 1 ? (size_t)E * 8 * nsp : 0);
    std::fclose(f);
    // the index structures en234gpu_create builds (en234gpu.cu: adjacency, block-CSR pattern, slot map)
    std::vector<double> xs(N), ys(N), zs(N);
    for (int n = 0; n < N; ++n) { xs[n] = coords[3 * n]; ys[n] = coords[3 * n + 1]; zs[n] = coords[3 * n + 2]; }
    std::vector<int> conn((size_t)8 * E), adj_ptr(N + 1, 0);
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < 8; ++a) { const int n = connectivity[8 * (size_t)e + a] - 1; conn[(size_t)a * E + e] = n; adj_ptr[n + 1]++; }
    for (int n = 0; n < N; ++n) adj_ptr[n + 1] += adj_ptr[n];
    std::vector<int> adj((size_t)8 * E), fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < 8; ++a) adj[fill[conn[(size_t)a * E + e]]++] = e * 8 + a;
    std::vector<int> browptr(N + 1, 0), bcol, cols;
    std::vector<uint8_t> slot((size_t)64 * E);
    for (int i = 0; i < N; ++i) {
        cols.clear();
        for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; ++k)
            for (int b = 0; b < 8; ++b) cols.push_back(conn[(size_t)b * E + (adj[k] >> 3)]);
        if (cols.empty()) cols.push_back(i);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; ++k)
            for (int b = 0; b < 8; ++b)
                slot[(size_t)k * 8 + b] = (uint8_t)(std::lower_bound(cols.begin(), cols.end(), conn[(size_t)b * E + (adj[k] >> 3)]) - cols.begin());
        bcol.insert(bcol.end(), cols.begin(), cols.end());
        browptr[i + 1] = (int)bcol.size();
    }
    std::vector<int> mat(E, 0);
    Material m0{};
    m0.kind = 0; m0.d11 = props[0]; m0.d12 = props[1]; m0.d44 = props[2];
    Mesh M;
    M.N = N; M.E = E; M.x = xs.data(); M.y = ys.data(); M.z = zs.data(); M.conn = conn.data(); M.mat = mat.data(); M.mats = &m0;
    M.n_mats = 1; M.adj_ptr = adj_ptr.data(); M.adj = adj.data(); M.adj_pos = nullptr; M.slot = slot.data();
    M.browptr = browptr.data(); M.bcol = bcol.data();

    std::vector<double> F((size_t)std::max(nf, 1) * N), val(9 * bcol.size()), minv(3 * (size_t)N), lump(N), mass(3 * (size_t)N);
    ProjArgs A{};
    A.M = M; A.ut = ut.data(); A.du = du.data(); A.sv = source == 1 ? sv.data() : nullptr; A.nsp = nsp; A.source = source; A.nf = nf;
    for (int k = 0; k < nf; ++k) A.code[k] = codes[k];
    A.F = F.data();
    MassArgs Ma{};
    Ma.M = M; Ma.val = val.data(); Ma.minv = minv.data(); Ma.lump = lump.data(); Ma.raw_diag = 1; Ma.lump_stride = 1;
    const int grid = (N + PROJ_THREADS - 1) / PROJ_THREADS;
    for (int b = 0; b < grid; ++b)
        for (int t = 0; t < PROJ_THREADS; ++t) {
            blockIdx.x = b; threadIdx.x = t;
            k_project_rhs(A);
            k_project_mass(Ma);
            k_dyn_lumped_mass(M, density.data(), mass.data());
        }
    std::vector<double> mdiag(N), rowsum(N, 0.0);
    for (int n = 0; n < N; ++n) {
        mdiag[n] = minv[3 * (size_t)n];
        const int nb = browptr[n + 1] - browptr[n], L = 3 * nb;
        const double* row = val.data() + 9 * (size_t)browptr[n];
        for (int k = 0; k < nb; ++k) {
            rowsum[n] += row[3 * k];
            // diag(m, m, m) blocks: the three diagonal entries are equal, everything else is zero
            if (row[3 * k] != row[L + 3 * k + 1] || row[3 * k] != row[2 * L + 3 * k + 2] || row[3 * k + 1] != 0.0 || row[L + 3 * k] != 0.0) return 3;
        }
    }
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    std::fwrite(F.data(), sizeof(double), (size_t)nf * N, o);
    std::fwrite(mdiag.data(), sizeof(double), N, o);
    std::fwrite(lump.data(), sizeof(double), N, o);
    std::fwrite(rowsum.data(), sizeof(double), N, o);
    std::fwrite(mass.data(), sizeof(double), 3 * (size_t)N, o);
    std::fclose(o);
    return 0;
}
