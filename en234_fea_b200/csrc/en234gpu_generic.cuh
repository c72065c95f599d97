// The other 3-D shapes of element 1001: 4-node and 10-node tetrahedra, 20-node bricks (el_linelast_3Dbasic.f90:82-85 picks
// 1 / 4 / 27 integration points for them; rules and shape functions: Element_Utilities.f90:586-689, :863-1025).
// SURVEY section 8(f) row 3.  These meshes go through a plain, general path — any number of nodes per element, any row
// length — that reuses everything downstream of the assembly unchanged (boundary conditions, block-CSR SpMV with the long-row
// loop, PCG): one thread per 3x3 block of an element stiffness (the thread integrates its block over the element's points),
// one thread per (element, node) for the residual, one warp per node row for the ordered gather.  Not a tuned path: the
// hex8 kernels of en234gpu_elem.cuh are the benchmarked ones.
// The shape-function table is the reference's, INCLUDING three entries that are not the derivatives of its own shape
// functions (tet10 df(10,1), hex20 df(11,1:2) and the sign of df(15,1)): results must equal the reference's on the same
// input; the oracle's copy of the table is pinned bit for bit against the Fortran text (tests/test_oracle_cpu.py).
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

struct GenMesh {
    int N, E, npe;
    const double *x, *y, *z;
    const int* conn;            // conn[e * npe + a], 0-based nodes
    const int* mat;
    const Material* mats;
    const int* adj_ptr;         // node -> adjacency entries, ascending element
    const int* adj;             // e * 32 + a
    const uint8_t* slot;        // npe per adjacency entry: position of column conn[e][b] in the node's block row
    const int* browptr;
    const int* bcol;
};

// initialize_integration_points, 3-D branch (Element_Utilities.f90:586-689): literals without a D exponent are single
// precision in the reference and are reproduced as (double)(float) values
__device__ __forceinline__ int gen_points(int nn) { return nn == 4 ? 1 : nn == 10 ? 4 : nn == 8 ? 8 : 27; }
__device__ __forceinline__ void gen_point(int nn, int k, double* xi, double& w) {
    if (nn == 4) { xi[0] = xi[1] = xi[2] = 0.25; w = 1.0 / 6.0; }
    else if (nn == 10) {
        const double a = (double)0.58541020f, b = (double)0.13819660f;
        xi[0] = k == 0 ? a : b; xi[1] = k == 1 ? a : b; xi[2] = k == 2 ? a : b;
        w = 1.0 / 24.0;
    } else if (nn == 8) {
        const double g = (double)0.5773502692f;
        xi[0] = (k & 1) ? g : -g; xi[1] = (k & 2) ? g : -g; xi[2] = (k & 4) ? g : -g;
        w = 1.0;
    } else {
        const double g = (double)0.7745966692f;
        const int i = k % 3, j = (k / 3) % 3, l = k / 9;
        const double x1[3] = {-g, 0.0, g}, w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
        xi[0] = x1[i]; xi[1] = x1[j]; xi[2] = x1[l];
        w = w1[i] * w1[j] * w1[l];
    }
}

// calculate_shapefunctions, derivatives only (Element_Utilities.f90:863-1025), as written there
__device__ void gen_dshape(int nn, const double* xi, double (*df)[3]) {
    for (int a = 0; a < nn; ++a) df[a][0] = df[a][1] = df[a][2] = 0.0;
    const double x = xi[0], y = xi[1], z = xi[2];
    if (nn == 4) {
        df[0][0] = 1.; df[1][1] = 1.; df[2][2] = 1.;
        df[3][0] = -1.; df[3][1] = -1.; df[3][2] = -1.;
    } else if (nn == 10) {
        const double x4 = 1. - x - y - z;
        df[0][0] = (4. * x - 1.); df[1][1] = (4. * y - 1.); df[2][2] = (4. * z - 1.);
        df[3][0] = -(4. * x4 - 1.); df[3][1] = -(4. * x4 - 1.); df[3][2] = -(4. * x4 - 1.);
        df[4][0] = 4. * y; df[4][1] = 4. * x;
        df[5][1] = 4. * z; df[5][2] = 4. * y;
        df[6][0] = 4. * z; df[6][2] = 4. * x;
        df[7][0] = 4. * (x4 - x); df[7][1] = -4. * x; df[7][2] = -4. * x;
        df[8][0] = -4. * y; df[8][1] = 4. * (x4 - y); df[8][2] = -4. * y;
        df[9][0] = -4. * z * x4;                           // sic (:902)
        df[9][1] = -4. * z; df[9][2] = 4. * (x4 - z);
    } else {
        const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        if (nn == 8) {
            for (int a = 0; a < 8; ++a) {
                const double fx = 1. + sx[a] * x, fy = 1. + sy[a] * y, fz = 1. + sz[a] * z;
                df[a][0] = sx[a] * (fy * fz / 8.); df[a][1] = sy[a] * (fx * fz / 8.); df[a][2] = sz[a] * (fx * fy / 8.);
            }
            return;
        }
        for (int a = 0; a < 8; ++a) {                      // corner nodes (:946-973)
            const double fx = 1. + sx[a] * x, fy = 1. + sy[a] * y, fz = 1. + sz[a] * z;
            const double g = sx[a] * x + sy[a] * y + sz[a] * z - 2.;
            df[a][0] = (sx[a] * (fy * fz * g) + sx[a] * (fx * fy * fz)) / 8.;
            df[a][1] = (sy[a] * (fx * fz * g) + sy[a] * (fx * fy * fz)) / 8.;
            df[a][2] = (sz[a] * (fx * fy * g) + sz[a] * (fx * fy * fz)) / 8.;
        }
        const double xm = 1. - x, xp = 1. + x, ym = 1. - y, yp = 1. + y, zm = 1. - z, zp = 1. + z;
        const double x2 = 1. - x * x, y2 = 1. - y * y, z2 = 1. - z * z;
        df[8][0] = -2. * x * ym * zm / 4.;  df[8][1] = -x2 * zm / 4.;            df[8][2] = -x2 * ym / 4.;
        df[9][0] = y2 * zm / 4.;            df[9][1] = -2. * y * xp * zm / 4.;   df[9][2] = -y2 * xp / 4.;
        df[10][0] = -2. * x * ym * zm / 4.; df[10][1] = -x2 * zm / 4.;           df[10][2] = -x2 * ym / 4.;   // sic (:991-993)
        df[11][0] = -y2 * zm / 4.;          df[11][1] = -2. * y * xm * zm / 4.;  df[11][2] = -y2 * xm / 4.;
        df[12][0] = -2. * x * ym * zp / 4.; df[12][1] = -x2 * zp / 4.;           df[12][2] = x2 * ym / 4.;
        df[13][0] = y2 * zp / 4.;           df[13][1] = -2. * y * xp * zp / 4.;  df[13][2] = y2 * xp / 4.;
        df[14][0] = 2. * x * yp * zp / 4.;  df[14][1] = x2 * zp / 4.;            df[14][2] = x2 * yp / 4.;    // sic (:1003)
        df[15][0] = -y2 * zp / 4.;          df[15][1] = -2. * y * xm * zp / 4.;  df[15][2] = y2 * xm / 4.;
        df[16][0] = -ym * z2 / 4.;          df[16][1] = -xm * z2 / 4.;           df[16][2] = -z * xm * ym / 2.;
        df[17][0] = ym * z2 / 4.;           df[17][1] = -xp * z2 / 4.;           df[17][2] = -z * xp * ym / 2.;
        df[18][0] = yp * z2 / 4.;           df[18][1] = xp * z2 / 4.;            df[18][2] = -z * xp * yp / 2.;
        df[19][0] = -yp * z2 / 4.;          df[19][1] = xm * z2 / 4.;            df[19][2] = -z * xm * yp / 2.;
    }
}

// dN/dx of all nodes of element e at integration point k; returns w |J| (0 flags a singular Jacobian)
__device__ double gen_gradients(const GenMesh& M, int e, int k, double (*dNdx)[3], bool& bad) {
    double xi[3], w, df[20][3], J[9], xc[20][3];
    gen_point(M.npe, k, xi, w);
    gen_dshape(M.npe, xi, df);
    for (int a = 0; a < M.npe; ++a) {
        const int n = M.conn[(size_t)e * M.npe + a];
        xc[a][0] = M.x[n]; xc[a][1] = M.y[n]; xc[a][2] = M.z[n];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double t = 0.0;
            for (int a = 0; a < M.npe; ++a) t += xc[a][i] * df[a][j];
            J[3 * i + j] = t;
        }
    Jac jc;
    invert3(J, jc);
    if (!(jc.det != 0.0)) bad = true;
    for (int a = 0; a < M.npe; ++a)
        for (int j = 0; j < 3; ++j) dNdx[a][j] = df[a][0] * jc.Ji[j] + df[a][1] * jc.Ji[3 + j] + df[a][2] * jc.Ji[6 + j];
    return w * jc.det;
}

// K_e block (a, b) = sum_k w|J| (lam S + mu S^T + mu tr(S) I), S = grad N_a (x) grad N_b  — the integrand of B^T D B
// (el_linelast_3Dbasic.f90:130-131) for the isotropic D; one thread per block.  Ke[(e npe^2 + a npe + b) 9 + 3 r + c]
__global__ void __launch_bounds__(128) k_gen_blocks(GenMesh M, double* __restrict__ Ke, CgScalars* S) {
    const long long total = (long long)M.E * M.npe * M.npe;
    bool bad = false;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / (M.npe * M.npe)), ab = (int)(t % (M.npe * M.npe)), a = ab / M.npe, b = ab % M.npe;
        const Material mat = M.mats[M.mat[e]];
        double Sm[9];
        for (int q = 0; q < 9; ++q) Sm[q] = 0.0;
        const int npts = gen_points(M.npe);
        for (int k = 0; k < npts; ++k) {
            double dNdx[20][3];
            const double wd = gen_gradients(M, e, k, dNdx, bad);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) Sm[3 * r + c] += (wd * dNdx[a][r]) * dNdx[b][c];
        }
        const double lam = mat.d12, mu = mat.d44, tr = mu * (Sm[0] + Sm[4] + Sm[8]);
        double* out = Ke + 9 * (size_t)t;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) out[3 * r + c] = lam * Sm[3 * r + c] + mu * Sm[3 * c + r] + (r == c ? tr : 0.0);
    }
    if (bad && S) atomicOr(&S->nanflag, 1);
}

// R_e rows of local node a: - sum_k w|J| sigma . grad N_a (el_linelast_3Dbasic.f90:124-128); Re[(e npe + a) 3 + i]
__global__ void __launch_bounds__(128) k_gen_residual(GenMesh M, const double* __restrict__ ut, const double* __restrict__ du,
                                                      double* __restrict__ Re, CgScalars* S) {
    const long long total = (long long)M.E * M.npe;
    bool bad = false;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / M.npe), a = (int)(t % M.npe);
        const Material mat = M.mats[M.mat[e]];
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        const int npts = gen_points(M.npe);
        for (int k = 0; k < npts; ++k) {
            double dNdx[20][3], H[9], s[9];
            const double wd = gen_gradients(M, e, k, dNdx, bad);
            for (int q = 0; q < 9; ++q) H[q] = 0.0;
            for (int c = 0; c < M.npe; ++c) {
                const int n = M.conn[(size_t)e * M.npe + c];
                for (int i = 0; i < 3; ++i) {
                    const double u = ut[3 * (size_t)n + i] + du[3 * (size_t)n + i];
                    for (int j = 0; j < 3; ++j) H[3 * i + j] += u * dNdx[c][j];
                }
            }
            stress_linear(mat, H, s);
            r0 -= (s[0] * dNdx[a][0] + s[1] * dNdx[a][1] + s[2] * dNdx[a][2]) * wd;
            r1 -= (s[3] * dNdx[a][0] + s[4] * dNdx[a][1] + s[5] * dNdx[a][2]) * wd;
            r2 -= (s[6] * dNdx[a][0] + s[7] * dNdx[a][1] + s[8] * dNdx[a][2]) * wd;
        }
        Re[3 * (size_t)t] = r0; Re[3 * (size_t)t + 1] = r1; Re[3 * (size_t)t + 2] = r2;
    }
    if (bad && S) atomicOr(&S->nanflag, 1);
}

struct GenGatherArgs {
    GenMesh M;
    const double* Ke;
    const double* Re;
    double *val, *rhs, *rforce, *partials, *diag;
    CgScalars* S;
};

// One warp per node row: the row of the block-CSR matrix is zeroed and the node's (element, local node) contributions are
// added one after the other in ascending element order (the reference's serial order, solver_direct.f90:264-287); within a
// contribution the blocks go to different columns, so the lanes never meet.  rhs, rforce = -rhs, ||rhs||^2, NaN scan and
// the assembled diagonal as in k_gather_rows.
__global__ void __launch_bounds__(THREADS) k_gen_gather(GenGatherArgs A) {
    __shared__ double red[WARPS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GenMesh& M = A.M;
    const int npe = M.npe, nd2 = 9 * npe * npe;
    double norm2 = 0.0;
    bool bad = false;
    for (int i = blockIdx.x * WARPS + warp; i < M.N; i += gridDim.x * WARPS) {
        const int rb = M.browptr[i], nb = M.browptr[i + 1] - rb, L = 3 * nb;
        const int k0 = M.adj_ptr[i], k1 = M.adj_ptr[i + 1];
        double* row = A.val + 9 * (size_t)rb;
        for (int q = lane; q < 9 * nb; q += 32) row[q] = 0.0;
        __syncwarp();
        double R = 0.0;
        int dslot = 0;
        for (int k = k0; k < k1; ++k) {
            const int e = M.adj[k] >> 5, a = M.adj[k] & 31;
            if (k == k0) dslot = M.slot[(size_t)k * npe + a];
            if (lane < 3) R += A.Re[3 * ((size_t)e * npe + a) + lane];
            const double* src = A.Ke + (size_t)e * nd2 + (size_t)a * npe * 9;
            for (int q = lane; q < 9 * npe; q += 32) {
                const int b = q / 9, rc = q % 9;
                row[(rc / 3) * L + 3 * (int)M.slot[(size_t)k * npe + b] + rc % 3] += src[q];
            }
            __syncwarp();
        }
        if (lane < 3) {
            A.diag[3 * (size_t)i + lane] = k0 < k1 ? row[lane * L + 3 * dslot + lane] : 0.0;
            A.rhs[3 * (size_t)i + lane] = R;
            A.rforce[3 * (size_t)i + lane] = -R;
            norm2 += R * R;
            if (R != R || R == R + 1.0) bad = true;
        }
        __syncwarp();
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
    const double t = block_sum<THREADS>(norm2, red);
    if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
}

}  // namespace en234k
