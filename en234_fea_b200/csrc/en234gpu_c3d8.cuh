// Built-in C3D8 continuum element (element flag 10003) on the device: finite-strain hypoelastic B-bar hex8 calling the
// linear-elastic ABAQUS UMAT.  Reference: src/continuum_elements.f90:368-761 (continuum_element_static_3D, 8-node
// branch) and user_codes/src/abaqus_umat_elastic.for:253-308 (3-D branch).  SURVEY 8(f) row 1: the reference's only
// golden output (Output_Files/Abaqus_umat_holeplate_3d.out) runs this element, so the device path is checked against
// reference output directly.  One thread per element: the tangent is a dense UNSYMMETRIC 24x24 matrix per element with
// a two-pass structure (volume averages first), kept in thread-local memory; this is the plain first version of a
// "next" row, not a tuned kernel.  Output goes to the element scratch of the two-kernel assembly (en234gpu_elem.cuh), the
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

}  // namespace c3d8

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
        const double* si = A.sv0 + (size_t)e * C3D8_NSV + C3D8_NSV_POINT * kint;
        M3 st = {{{si[0], si[3], si[4]}, {si[3], si[1], si[5]}, {si[4], si[5], si[2]}}};
        st = mulT(mul(dR, st), dR);
        M3 en = {{{si[6], 0.5 * si[9], 0.5 * si[10]}, {0.5 * si[9], si[7], 0.5 * si[11]}, {0.5 * si[10], 0.5 * si[11], si[8]}}};
        en = mulT(mul(dR, en), dR);
        double stress[6] = {st.a[0][0], st.a[1][1], st.a[2][2], st.a[0][1], st.a[0][2], st.a[1][2]};
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
        double* so = A.sv1 + (size_t)e * C3D8_NSV + C3D8_NSV_POINT * kint;      // :718-722
        for (int i = 0; i < 6; ++i) { so[i] = stress[i]; so[6 + i] = strain[i] + dstran[i]; }
        so[12] = si[12]; so[13] = si[13]; so[14] = si[14];
        // tangent pieces (:648-716)
        M3 H, Lam;
        {
            const M3 FH = mul(F1, Fmi);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) { H.a[i][j] = FH.a[j][i]; Lam.a[i][j] = (i == j ? 1.0 : 0.0) - 0.5 * L.a[i][j]; }
        }
        double Mx[6][9];                                              // Dcorot depsdf + Dspin
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
