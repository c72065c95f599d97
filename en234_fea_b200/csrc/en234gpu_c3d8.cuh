// Built-in C3D8 continuum element (element flag 10003) on the device: finite-strain hypoelastic B-bar hex8 calling the
// linear-elastic ABAQUS UMAT.  Reference: src/continuum_elements.f90:368-761 (continuum_element_static_3D, 8-node
// branch) and user_codes/src/abaqus_umat_elastic.for:253-308 (3-D branch).  SURVEY 8(f) row 1: the reference's only
// golden output (Output_Files/Abaqus_umat_holeplate_3d.out) runs this element, so the device path is checked against
// reference output directly.  The tangent is a dense UNSYMMETRIC 24x24 matrix per element with a two-pass structure
// (volume averages first).  Two kernels with the same arithmetic: k_element_c3d8_warp (default: one warp per element, 3 x 6
// register tiles, integration-point data in shared memory) and k_element_c3d8 (round 1: one thread per element, everything
// in thread-local memory; EN234_C3D8_TPE=1 selects it for A/B).  Output goes to the element scratch of the two-kernel assembly (en234gpu_elem.cuh), the
// deterministic row gather is shared with the user elements.
// State variables per integration point (continuum_elements.f90:718-722): stress(6), strain(6), 3 spare = 15.
#pragma once
#include "en234gpu_kernels.cuh"

namespace en234k {

constexpr int C3D8_NSV_POINT = 15;
constexpr int C3D8_NSV = 8 * C3D8_NSV_POINT;

struct C3d8Args {
    Mesh M;
    const double* ut;       // dof_total
    const double* du;       // dof_increment
    const double* sv0;      // initial_state_variables  [e][120]
    double* sv1;            // updated_state_variables
    double* Ke;             // element scratch (layout of en234gpu_elem.cuh), null: residual only
    double* Re;
    CgScalars* S;
};

namespace c3d8 {

struct M3 { double a[3][3]; };

__device__ inline M3 mul(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C.a[i][j] = A.a[i][0] * B.a[0][j] + A.a[i][1] * B.a[1][j] + A.a[i][2] * B.a[2][j];
    return C;
}
__device__ inline M3 mulT(const M3& A, const M3& B) {      // A B^T
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C.a[i][j] = A.a[i][0] * B.a[j][0] + A.a[i][1] * B.a[j][1] + A.a[i][2] * B.a[j][2];
    return C;
}
// determinant and cofactor inverse in the operation order of Element_Utilities.f90:51-68, :99-123
__device__ inline double det(const M3& A) {
    return A.a[0][0] * A.a[1][1] * A.a[2][2] - A.a[0][0] * A.a[1][2] * A.a[2][1] - A.a[0][1] * A.a[1][0] * A.a[2][2] +
           A.a[0][1] * A.a[1][2] * A.a[2][0] + A.a[0][2] * A.a[1][0] * A.a[2][1] - A.a[0][2] * A.a[1][1] * A.a[2][0];
}
__device__ inline double inverse(const M3& A, M3& Ai) {
    const double d = det(A);
    Ai.a[0][0] = (A.a[1][1] * A.a[2][2] - A.a[1][2] * A.a[2][1]) / d;
    Ai.a[1][0] = -(A.a[1][0] * A.a[2][2] - A.a[1][2] * A.a[2][0]) / d;
    Ai.a[2][0] = (A.a[1][0] * A.a[2][1] - A.a[1][1] * A.a[2][0]) / d;
    Ai.a[0][1] = -(A.a[0][1] * A.a[2][2] - A.a[0][2] * A.a[2][1]) / d;
    Ai.a[1][1] = (A.a[0][0] * A.a[2][2] - A.a[0][2] * A.a[2][0]) / d;
    Ai.a[2][1] = -(A.a[0][0] * A.a[2][1] - A.a[0][1] * A.a[2][0]) / d;
    Ai.a[0][2] = (A.a[0][1] * A.a[1][2] - A.a[0][2] * A.a[1][1]) / d;
    Ai.a[1][2] = -(A.a[0][0] * A.a[1][2] - A.a[0][2] * A.a[1][0]) / d;
    Ai.a[2][2] = (A.a[0][0] * A.a[1][1] - A.a[0][1] * A.a[1][0]) / d;
    return d;
}

// kinematics of one integration point (continuum_elements.f90:519-545 / :562-577)
struct Point {
    double dNdy[8][3];      // spatial shape function gradients dN/dx F1^-1
    M3 F0, F1;              // deformation gradient at the start / end of the increment
    double J0, J1, wdet;    // det F0, det F1, w |dx/dxi|
};

__device__ inline void point(const double* xc, const double* u0, const double* du, int kint, Point& p) {
    double fp[3], fm[3];
    for (int j = 0; j < 3; ++j) {
        const double xi = ((kint >> j) & 1) ? EN234_GP : -EN234_GP;
        fp[j] = 1.0 + xi; fm[j] = 1.0 - xi;
    }
    double dN[8][3];
    for (int a = 0; a < 8; ++a) {
        const int s0 = sgn(a, 0), s1 = sgn(a, 1), s2 = sgn(a, 2);
        const double f0 = s0 > 0 ? fp[0] : fm[0], f1 = s1 > 0 ? fp[1] : fm[1], f2 = s2 > 0 ? fp[2] : fm[2];
        dN[a][0] = s0 * (f1 * f2 * 0.125); dN[a][1] = s1 * (f0 * f2 * 0.125); dN[a][2] = s2 * (f0 * f1 * 0.125);
    }
    M3 J, Ji;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int a = 0; a < 8; ++a) s += xc[3 * a + i] * dN[a][j];
            J.a[i][j] = s;
        }
    p.wdet = inverse(J, Ji);                                   // w = 1
    double dNdx[8][3];
    for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 3; ++j) dNdx[a][j] = dN[a][0] * Ji.a[0][j] + dN[a][1] * Ji.a[1][j] + dN[a][2] * Ji.a[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s0 = 0.0, s1 = 0.0;
            for (int a = 0; a < 8; ++a) { s0 += u0[3 * a + i] * dNdx[a][j]; s1 += du[3 * a + i] * dNdx[a][j]; }
            p.F0.a[i][j] = s0 + (i == j ? 1.0 : 0.0);
            p.F1.a[i][j] = p.F0.a[i][j] + s1;
        }
    M3 F1i;
    p.J1 = inverse(p.F1, F1i);
    p.J0 = det(p.F0);
    for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 3; ++j) p.dNdy[a][j] = dNdx[a][0] * F1i.a[0][j] + dNdx[a][1] * F1i.a[1][j] + dNdx[a][2] * F1i.a[2][j];
}

// d(strain-like 6-vector)/d(9-vector) operator of continuum_elements.f90:652-669:
// rows (11,22,33,12,13,23), columns (11,22,33,12,21,13,31,23,32); row ii: A(i,k)B(i,l), row ij: A(i,k)B(j,l)+A(j,k)B(i,l)
__device__ inline void op69(const M3& A, const M3& B, double out[6][9]) {
    const int RI[6] = {0, 1, 2, 0, 0, 1}, RJ[6] = {0, 1, 2, 1, 2, 2};
    const int CK[9] = {0, 1, 2, 0, 1, 0, 2, 1, 2}, CL[9] = {0, 1, 2, 1, 0, 2, 0, 2, 1};
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 9; ++c) {
            const int i = RI[r], j = RJ[r], k = CK[c], l = CL[c];
            out[r][c] = (i == j) ? A.a[i][k] * B.a[i][l] : A.a[i][k] * B.a[j][l] + A.a[j][k] * B.a[i][l];
        }
}

// Stress update and tangent operators of one integration point (continuum_elements.f90:559-716): volumetric scaling of
// F0 / F1 to the element averages, Hughes-Winget rotation of the stored stress / strain, the elastic UMAT
// (abaqus_umat_elastic.for:284-305), the updated state variables, and Mx = D_corot d(eps)/dF + D_spin (6 x 9).
// si / so: the point's 15 state variables at the start / end of the increment.
__device__ inline void point_material(const Point& p, double J0bar, double J1bar, double E, double xnu, const double* si,
                                      double* so, double* stress /*6*/, double (*Mx)[9]) {
        const double s0 = pow(J0bar / p.J0, 1.0 / 3.0), s1 = pow(J1bar / p.J1, 1.0 / 3.0);
        M3 F0, F1, dF, Fmid, Fmi;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                F0.a[i][j] = p.F0.a[i][j] * s0;
                F1.a[i][j] = p.F1.a[i][j] * s1;
                dF.a[i][j] = F1.a[i][j] - F0.a[i][j];
                Fmid.a[i][j] = 0.5 * (F1.a[i][j] + F0.a[i][j]);
            }
        inverse(Fmid, Fmi);
        const M3 L = mul(dF, Fmi);
        M3 deps, Aw, Bw, Bwi;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double dW = 0.5 * (L.a[i][j] - L.a[j][i]);
                deps.a[i][j] = 0.5 * (L.a[i][j] + L.a[j][i]);
                Aw.a[i][j] = (i == j ? 1.0 : 0.0) + 0.5 * dW;         // Hughes-Winget rotation increment :591-595
                Bw.a[i][j] = (i == j ? 1.0 : 0.0) - 0.5 * dW;
            }
        inverse(Bw, Bwi);
        const M3 dR = mul(Bwi, Aw);
        M3 st = {{{si[0], si[3], si[4]}, {si[3], si[1], si[5]}, {si[4], si[5], si[2]}}};
        st = mulT(mul(dR, st), dR);
        M3 en = {{{si[6], 0.5 * si[9], 0.5 * si[10]}, {0.5 * si[9], si[7], 0.5 * si[11]}, {0.5 * si[10], 0.5 * si[11], si[8]}}};
        en = mulT(mul(dR, en), dR);
        stress[0] = st.a[0][0]; stress[1] = st.a[1][1]; stress[2] = st.a[2][2]; stress[3] = st.a[0][1]; stress[4] = st.a[0][2]; stress[5] = st.a[1][2];
        const double strain[6] = {en.a[0][0], en.a[1][1], en.a[2][2], en.a[0][1], en.a[0][2], en.a[1][2]};
        const double dstran[6] = {deps.a[0][0], deps.a[1][1], deps.a[2][2], 2.0 * deps.a[0][1], 2.0 * deps.a[0][2], 2.0 * deps.a[1][2]};
        // UMAT: ddsdde and stress += ddsdde dstran (abaqus_umat_elastic.for:284-305)
        double D[6][6];
        {
            const double f = E / ((1.0 + xnu) * (1.0 - 2.0 * xnu));
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) D[i][j] = 0.0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) D[i][j] = ((i == j) ? 1.0 - xnu : xnu) * f;
            D[3][3] = D[4][4] = D[5][5] = 0.5 * (1.0 - 2.0 * xnu) * f;
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) stress[i] = stress[i] + D[i][j] * dstran[j];
        }
        for (int i = 0; i < 6; ++i) { so[i] = stress[i]; so[6 + i] = strain[i] + dstran[i]; }
        so[12] = si[12]; so[13] = si[13]; so[14] = si[14];
        // tangent pieces (:648-716)
        M3 H, Lam;
        {
            const M3 FH = mul(F1, Fmi);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) { H.a[i][j] = FH.a[j][i]; Lam.a[i][j] = (i == j ? 1.0 : 0.0) - 0.5 * L.a[i][j]; }
        }
        // Mx: Dcorot depsdf + Dspin
        {
            double depsdf[6][9], w1[6][9], w2[6][9];
            op69(Lam, H, depsdf);
            const M3 S1 = {{{stress[0], stress[3], stress[4]}, {stress[3], stress[1], stress[5]}, {stress[4], stress[5], stress[2]}}};
            op69(Lam, mul(S1, H), w1);
            op69(mul(S1, Lam), H, w2);
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 9; ++c) {
                    double s = 0.0;
                    for (int k = 0; k < 6; ++k) s += D[r][k] * depsdf[k][c];
                    Mx[r][c] = s + 0.5 * (w1[r][c] - w2[r][c]);
                }
        }
}

}  // namespace c3d8

// ---- one warp per element ----------------------------------------------------------------------------------------------
// Same arithmetic, in the same order per value, as k_element_c3d8 below (which stays as the A/B check, EN234_C3D8_TPE=1), but
// without the 10 kB of thread-local arrays: a warp takes GROUP elements per pass so that GROUP x 8 (element, integration point) pairs share the lanes
// in the point phases (kinematics, stress update, Mx), which leave their results in shared memory; then per element the 576 entries of Pbar and of the tangent are split into 3 x 6 register tiles (one per lane) that
// stay in registers over the integration-point loop; B-bar, B-star, Sn and G = Mx B-star of the current point are built cooperatively in shared memory.
#ifndef EN234_C3D8_WARPS
#define EN234_C3D8_WARPS 3            // measured on 64^3 C3D8 elements (assembly incl. gather): (warps, group, min CTAs/SM)
#define EN234_C3D8_GROUP 2            //   (4,1,3) 7.6 ms, (2,2,5) 8.6 ms, (3,2,4) 6.95 ms, (2,4,3) 8.3 ms; thread per element 23.9 ms
#define EN234_C3D8_MINB 4
#endif
constexpr int C3D8_WARPS = EN234_C3D8_WARPS;
constexpr int C3D8_GROUP = EN234_C3D8_GROUP;   // elements per warp pass (1, 2, 4): GROUP x 8 integration points share the lanes in the point phases

struct C3d8Warp {
    double xin[C3D8_GROUP][72];             // coordinates, dof_total, dof_increment of the 8 nodes
    c3d8::Point pt[C3D8_GROUP][8];
    double stress[C3D8_GROUP][8][6];
    double Mx[C3D8_GROUP][8][6][9];
    double dNbar[24];
    double Bb[6][24], Bs[9][24], G[6][24], Sn[24];
};

__global__ void __launch_bounds__(32 * C3D8_WARPS, EN234_C3D8_MINB) k_element_c3d8_warp(C3d8Args A) {
    using namespace c3d8;
    extern __shared__ __align__(16) unsigned char c3d8_smem[];
    const Mesh& M = A.M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    C3d8Warp& W = reinterpret_cast<C3d8Warp*>(c3d8_smem)[warp];
    bool bad = false;
    for (int e0 = C3D8_GROUP * (blockIdx.x * C3D8_WARPS + warp); e0 < M.E; e0 += C3D8_GROUP * gridDim.x * C3D8_WARPS) {
        __syncwarp();
        for (int q = lane; q < 24 * C3D8_GROUP; q += 32) {
            const int el = q / 24, r = q % 24, a = r / 3, i = r % 3, e = e0 + el;
            if (e < M.E) {
                const int n = M.conn[(size_t)a * M.E + e];
                W.xin[el][r] = i == 0 ? M.x[n] : (i == 1 ? M.y[n] : M.z[n]);
                W.xin[el][24 + r] = A.ut[3 * (size_t)n + i];
                W.xin[el][48 + r] = A.du[3 * (size_t)n + i];
            }
        }
        __syncwarp();
        // ---- point phases: lane = (element of the group, integration point)
        const int pel = lane >> 3, pk = lane & 7, pe = e0 + pel;
        const bool pon = pel < C3D8_GROUP && pe < M.E;
        if (pon) {
            Point& p = W.pt[pel][pk];
            point(W.xin[pel], W.xin[pel] + 24, W.xin[pel] + 48, pk, p);
            if (!(p.wdet != 0.0) || !(p.J1 > 0.0)) bad = true;
        }
        __syncwarp();
        if (pon) {                         // volume averages of the lane's element, then its point's stress update (:559-716)
            double J0bar = 0.0, J1bar = 0.0, vol = 0.0;
            for (int k = 0; k < 8; ++k) {
                const double wd = W.pt[pel][k].wdet;
                vol += wd;
                J0bar += W.pt[pel][k].J0 * wd;
                J1bar += W.pt[pel][k].J1 * wd;
            }
            J0bar /= vol;
            J1bar /= vol;
            const Material mat = M.mats[M.mat[pe]];
            const double* si = A.sv0 + (size_t)pe * C3D8_NSV + C3D8_NSV_POINT * pk;
            double* so = A.sv1 + (size_t)pe * C3D8_NSV + C3D8_NSV_POINT * pk;
            double st[6], Mx[6][9];
            point_material(W.pt[pel][pk], J0bar, J1bar, mat.d11, mat.d12, si, so, st, Mx);       // kind 2 stores (E, nu) there
            for (int r = 0; r < 6; ++r) {
                W.stress[pel][pk][r] = st[r];
                for (int c = 0; c < 9; ++c) W.Mx[pel][pk][r][c] = Mx[r][c];
            }
        }
        __syncwarp();
      for (int el = 0; el < C3D8_GROUP; ++el) {
        const int e = e0 + el;
        if (e >= M.E) break;
        // ---- pass 1: volume averages (:519-557)
        double J0bar = 0.0, J1bar = 0.0, vol = 0.0;
        for (int k = 0; k < 8; ++k) {
            const double wd = W.pt[el][k].wdet;
            vol += wd;
            J0bar += W.pt[el][k].J0 * wd;
            J1bar += W.pt[el][k].J1 * wd;
        }
        J0bar /= vol;
        J1bar /= vol;
        if (lane < 24) {
            double s = 0.0;
            for (int k = 0; k < 8; ++k) s += W.pt[el][k].J1 * W.pt[el][k].dNdy[lane / 3][lane % 3] * W.pt[el][k].wdet;
            W.dNbar[lane] = s / (J1bar * vol);
        }
        __syncwarp();
        // lane -> node a (3 tangent rows) and column group cg (6 columns = nodes 2cg, 2cg+1): an 3 x 6 register tile
        const int ta = lane >> 2, tc = 6 * (lane & 3);
        double Pb[18], K[18];
        {
            double acc[18];
#pragma unroll
            for (int m = 0; m < 18; ++m) acc[m] = 0.0;
            for (int k = 0; k < 8; ++k) {
                const Point& p = W.pt[el][k];
                const double da[3] = {p.dNdy[ta][0], p.dNdy[ta][1], p.dNdy[ta][2]};
#pragma unroll
                for (int jb = 0; jb < 2; ++jb) {
                    const double* db = p.dNdy[(tc / 3) + jb];
                    const double d0 = db[0], d1 = db[1], d2 = db[2];
#pragma unroll
                    for (int ri = 0; ri < 3; ++ri)
#pragma unroll
                        for (int kj = 0; kj < 3; ++kj) {
                            const double dbkj = kj == 0 ? d0 : (kj == 1 ? d1 : d2), dbri = ri == 0 ? d0 : (ri == 1 ? d1 : d2);
                            acc[ri * 6 + 3 * jb + kj] += p.J1 * (da[ri] * dbkj - da[kj] * dbri) * p.wdet;
                        }
                }
            }
#pragma unroll
            for (int m = 0; m < 18; ++m) {
                Pb[m] = acc[m] / (J1bar * vol) - W.dNbar[3 * ta + m / 6] * W.dNbar[tc + m % 6];
                K[m] = 0.0;
            }
        }
        // ---- pass 2: residual and tangent (:724-757)
        double R = 0.0;
        for (int k = 0; k < 8; ++k) {
            const Point& p = W.pt[el][k];
            const double* st = W.stress[el][k];
            double* bb = &W.Bb[0][0];
            double* bs = &W.Bs[0][0];
            for (int q = lane; q < 144; q += 32) bb[q] = 0.0;
            for (int q = lane; q < 216; q += 32) bs[q] = 0.0;
            __syncwarp();
            if (lane < 8) {
                const int a = lane;
                const double* d = p.dNdy[a];
                W.Bb[0][3 * a] = d[0]; W.Bb[1][3 * a + 1] = d[1]; W.Bb[2][3 * a + 2] = d[2];
                W.Bb[3][3 * a] = d[1]; W.Bb[3][3 * a + 1] = d[0];
                W.Bb[4][3 * a] = d[2]; W.Bb[4][3 * a + 2] = d[0];
                W.Bb[5][3 * a + 1] = d[2]; W.Bb[5][3 * a + 2] = d[1];
            }
            __syncwarp();
            if (lane < 24) {
                double s = 0.0;
                for (int r = 0; r < 6; ++r) s += W.Bb[r][lane] * st[r];
                W.Sn[lane] = s;
            }
            __syncwarp();
            for (int q = lane; q < 72; q += 32) {
                const int j = q % 24;
                bb[q] += (W.dNbar[j] - p.dNdy[j / 3][j % 3]) / 3.0;
                bs[q] = bb[q];
            }
            if (lane < 8) {
                const int a = lane;
                const double* d = p.dNdy[a];
                W.Bs[3][3 * a] = d[1]; W.Bs[4][3 * a + 1] = d[0]; W.Bs[5][3 * a] = d[2];
                W.Bs[6][3 * a + 2] = d[0]; W.Bs[7][3 * a + 1] = d[2]; W.Bs[8][3 * a + 2] = d[1];
            }
            __syncwarp();
            const double press = (st[0] + st[1] + st[2]) / 3.0;
            const double wdj = p.wdet * J1bar;
            if (lane < 24) {
                double s = 0.0;
                for (int r = 0; r < 6; ++r) s += W.Bb[r][lane] * st[r];
                R = R - s * wdj;
            }
            if (A.Ke) {
                for (int q = lane; q < 144; q += 32) {
                    const int r = q / 24, j = q % 24;
                    double s = 0.0;
                    for (int c = 0; c < 9; ++c) s += W.Mx[el][k][r][c] * W.Bs[c][j];
                    W.G[r][j] = s;
                }
                __syncwarp();
                double bbi[6][3];
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int ri = 0; ri < 3; ++ri) bbi[r][ri] = W.Bb[r][3 * ta + ri];
                const double da[3] = {p.dNdy[ta][0], p.dNdy[ta][1], p.dNdy[ta][2]};
#pragma unroll
                for (int jc = 0; jc < 6; ++jc) {
                    const int j = tc + jc, b = j / 3, kj = jc % 3;
                    double g[6];
#pragma unroll
                    for (int r = 0; r < 6; ++r) g[r] = W.G[r][j];
#pragma unroll
                    for (int ri = 0; ri < 3; ++ri) {
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < 6; ++r) s += bbi[r][ri] * g[r];
                        const double geo_s = da[kj] * W.Sn[3 * b + ri];
                        const double geo_p = da[kj] * p.dNdy[b][ri];
                        K[ri * 6 + jc] += (s - geo_s + press * (Pb[ri * 6 + jc] + geo_p)) * wdj;
                    }
                }
            }
            __syncwarp();
        }
        if (lane < 24) {
            A.Re[3 * (size_t)M.adj_pos[8 * (size_t)e + lane / 3] + lane % 3] = R;
            if (R != R) bad = true;
        }
        if (A.Ke) {
            double* base = A.Ke + (size_t)(e >> 1) * 1152;
#pragma unroll
            for (int m = 0; m < 18; ++m) {
                const int a = ta, r = m / 6, j = tc + m % 6, b = j / 3, c = j % 3;
                const int jj = ((a & 1) * 2 + (b & 1)) * 9 + 3 * r + c, l = (e & 1) * 16 + (a >> 1) * 4 + (b >> 1);
                base[(jj >> 1) * 64 + 2 * l + (jj & 1)] = K[m];
            }
        }
      }
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
}

__global__ void __launch_bounds__(64, 1) k_element_c3d8(C3d8Args A) {
    using namespace c3d8;
    const Mesh& M = A.M;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M.E) return;
    const Material mat = M.mats[M.mat[e]];
    double xc[24], u0[24], du[24];
    for (int a = 0; a < 8; ++a) {
        const int n = M.conn[(size_t)a * M.E + e];
        xc[3 * a] = M.x[n]; xc[3 * a + 1] = M.y[n]; xc[3 * a + 2] = M.z[n];
        for (int i = 0; i < 3; ++i) { u0[3 * a + i] = A.ut[3 * (size_t)n + i]; du[3 * a + i] = A.du[3 * (size_t)n + i]; }
    }
    double K[576], Pbar[576], R[24];
    for (int q = 0; q < 576; ++q) { K[q] = 0.0; Pbar[q] = 0.0; }
    for (int q = 0; q < 24; ++q) R[q] = 0.0;
    // ---- pass 1: volume averages (:519-557)
    double J0bar = 0.0, J1bar = 0.0, vol = 0.0, dNbar[24];
    for (int q = 0; q < 24; ++q) dNbar[q] = 0.0;
    Point p;
    bool bad = false;
    for (int kint = 0; kint < 8; ++kint) {
        point(xc, u0, du, kint, p);
        if (!(p.wdet != 0.0) || !(p.J1 > 0.0)) bad = true;
        const double wd = p.wdet;
        for (int a = 0; a < 8; ++a)
            for (int j = 0; j < 3; ++j) dNbar[3 * a + j] += p.J1 * p.dNdy[a][j] * wd;
        vol += wd;
        J0bar += p.J0 * wd;
        J1bar += p.J1 * wd;
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j)
                Pbar[i * 24 + j] += p.J1 * (p.dNdy[i / 3][i % 3] * p.dNdy[j / 3][j % 3] - p.dNdy[i / 3][j % 3] * p.dNdy[j / 3][i % 3]) * wd;
    }
    J0bar /= vol;
    J1bar /= vol;
    for (int q = 0; q < 24; ++q) dNbar[q] /= (J1bar * vol);
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) Pbar[i * 24 + j] = Pbar[i * 24 + j] / (J1bar * vol) - dNbar[i] * dNbar[j];
    // ---- pass 2: stress update, residual, tangent (:559-757)
    const double E = mat.d11, xnu = mat.d12;                        // kind 2 stores (E, nu) here
    for (int kint = 0; kint < 8; ++kint) {
        point(xc, u0, du, kint, p);
        const double* si = A.sv0 + (size_t)e * C3D8_NSV + C3D8_NSV_POINT * kint;
        double* so = A.sv1 + (size_t)e * C3D8_NSV + C3D8_NSV_POINT * kint;      // :718-722
        double stress[6], Mx[6][9];
        point_material(p, J0bar, J1bar, E, xnu, si, so, stress, Mx);
        // B-bar (6x24) and B-star (9x24) (:724-755)
        double Bb[6][24], Bs[9][24], Sn[24];
        for (int r = 0; r < 6; ++r)
            for (int j = 0; j < 24; ++j) Bb[r][j] = 0.0;
        for (int r = 0; r < 9; ++r)
            for (int j = 0; j < 24; ++j) Bs[r][j] = 0.0;
        for (int a = 0; a < 8; ++a) {
            const double* d = p.dNdy[a];
            Bb[0][3 * a] = d[0]; Bb[1][3 * a + 1] = d[1]; Bb[2][3 * a + 2] = d[2];
            Bb[3][3 * a] = d[1]; Bb[3][3 * a + 1] = d[0];
            Bb[4][3 * a] = d[2]; Bb[4][3 * a + 2] = d[0];
            Bb[5][3 * a + 1] = d[2]; Bb[5][3 * a + 2] = d[1];
        }
        for (int j = 0; j < 24; ++j) {
            double s = 0.0;
            for (int r = 0; r < 6; ++r) s += Bb[r][j] * stress[r];
            Sn[j] = s;
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 24; ++j) Bb[i][j] += (dNbar[j] - p.dNdy[j / 3][j % 3]) / 3.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 24; ++j) Bs[i][j] = Bb[i][j];
        for (int a = 0; a < 8; ++a) {
            const double* d = p.dNdy[a];
            Bs[3][3 * a] = d[1]; Bs[4][3 * a + 1] = d[0]; Bs[5][3 * a] = d[2];
            Bs[6][3 * a + 2] = d[0]; Bs[7][3 * a + 1] = d[2]; Bs[8][3 * a + 2] = d[1];
        }
        const double press = (stress[0] + stress[1] + stress[2]) / 3.0;
        const double wdj = p.wdet * J1bar;
        for (int i = 0; i < 24; ++i) {
            double s = 0.0;
            for (int r = 0; r < 6; ++r) s += Bb[r][i] * stress[r];
            R[i] = R[i] - s * wdj;
        }
        if (A.Ke) {
            double G[6][24];
            for (int r = 0; r < 6; ++r)
                for (int j = 0; j < 24; ++j) {
                    double s = 0.0;
                    for (int c = 0; c < 9; ++c) s += Mx[r][c] * Bs[c][j];
                    G[r][j] = s;
                }
            for (int i = 0; i < 24; ++i)
                for (int j = 0; j < 24; ++j) {
                    double s = 0.0;
                    for (int r = 0; r < 6; ++r) s += Bb[r][i] * G[r][j];
                    const int a = i / 3, ri = i % 3, b = j / 3, kj = j % 3;
                    const double geo_s = p.dNdy[a][kj] * Sn[3 * b + ri];
                    const double geo_p = p.dNdy[a][kj] * p.dNdy[b][ri];
                    K[i * 24 + j] += (s - geo_s + press * (Pbar[i * 24 + j] + geo_p)) * wdj;
                }
        }
    }
    for (int q = 0; q < 24; ++q) {
        A.Re[3 * (size_t)M.adj_pos[8 * (size_t)e + q / 3] + q % 3] = R[q];
        if (R[q] != R[q]) bad = true;
    }
    if (A.Ke) {
        double* base = A.Ke + (size_t)(e >> 1) * 1152;
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                const int a = i / 3, r = i % 3, b = j / 3, c = j % 3;
                const int jj = ((a & 1) * 2 + (b & 1)) * 9 + 3 * r + c, l = (e & 1) * 16 + (a >> 1) * 4 + (b >> 1);
                base[(jj >> 1) * 64 + 2 * l + (jj & 1)] = K[i * 24 + j];
            }
    }
    if (bad) atomicOr(&A.S->nanflag, 1);
}

}  // namespace en234k
