// ORACLE — test infrastructure only (see or_model.h).  SURVEY 8(f) rank-1 row: the built-in C3D8 B-bar finite-strain
// continuum element with the linear-elastic (hypoelastic) ABAQUS UMAT.  It is NOT on the GPU hot path; it is restated
// here only because the single golden output of the reference, Output_Files/Abaqus_umat_holeplate_3d.out, exercises it:
// reproducing that log pins the oracle's driver / assembly / BC / skyline-LDU / Newton logic with reference output.
//   element: src/continuum_elements.f90:368-761 (continuum_element_static_3D), hex8 branch
//   material: user_codes/src/abaqus_umat_elastic.for:253-308
#include <cmath>
#include <cstdio>
#include <cstring>

#include "or_model.h"

namespace oracle {

namespace {

typedef double M3[3][3];

const double SG[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
const double GPF = (double)0.5773502692f;       // Element_Utilities.f90:637-638

void mul(const M3 A, const M3 B, M3 C) {
    M3 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[i][k] * B[k][j]; T[i][j] = s; }
    std::memcpy(C, T, sizeof(M3));
}
void mulT(const M3 A, const M3 B, M3 C) {        // A * B^T
    M3 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[i][k] * B[j][k]; T[i][j] = s; }
    std::memcpy(C, T, sizeof(M3));
}
double det3(const M3 A) {                        // Element_Utilities.f90:51-68
    return A[0][0] * A[1][1] * A[2][2] - A[0][0] * A[1][2] * A[2][1] - A[0][1] * A[1][0] * A[2][2] +
           A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] - A[0][2] * A[1][1] * A[2][0];
}
double inv3(const M3 A, M3 Ai) {                 // cofactor inverse, Element_Utilities.f90:99-123
    const double d = det3(A);
    M3 C;
    C[0][0] = +(A[1][1] * A[2][2] - A[1][2] * A[2][1]);
    C[0][1] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]);
    C[0][2] = +(A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    C[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]);
    C[1][1] = +(A[0][0] * A[2][2] - A[0][2] * A[2][0]);
    C[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]);
    C[2][0] = +(A[0][1] * A[1][2] - A[0][2] * A[1][1]);
    C[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]);
    C[2][2] = +(A[0][0] * A[1][1] - A[0][1] * A[1][0]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ai[i][j] = C[j][i] / d;
    return d;
}

struct Point {
    double dNdx[8][3];     // reference-configuration gradients
    double dNdy[8][3];     // spatial gradients dNdx * F1^-1
    M3 F0, F1, dF;
    double J1, det;
};

void point(const double* xc, const double* du, const double* ut, int kint, Point& p) {
    double xi[3] = {(kint & 1) ? GPF : -GPF, (kint & 2) ? GPF : -GPF, (kint & 4) ? GPF : -GPF};
    double dN[8][3];
    for (int a = 0; a < 8; ++a) {
        const double f1 = 1.0 + SG[a][0] * xi[0], f2 = 1.0 + SG[a][1] * xi[1], f3 = 1.0 + SG[a][2] * xi[2];
        dN[a][0] = SG[a][0] * (f2 * f3 / 8.0);
        dN[a][1] = SG[a][1] * (f1 * f3 / 8.0);
        dN[a][2] = SG[a][2] * (f1 * f2 / 8.0);
    }
    M3 J, Ji;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { double s = 0; for (int a = 0; a < 8; ++a) s += xc[3 * a + i] * dN[a][j]; J[i][j] = s; }
    p.det = inv3(J, Ji);
    for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += dN[a][k] * Ji[k][j]; p.dNdx[a][j] = s; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s0 = 0, s1 = 0;
            for (int a = 0; a < 8; ++a) { s0 += ut[3 * a + i] * p.dNdx[a][j]; s1 += du[3 * a + i] * p.dNdx[a][j]; }
            p.F0[i][j] = s0 + (i == j ? 1.0 : 0.0);
            p.dF[i][j] = s1;
            p.F1[i][j] = p.F0[i][j] + s1;
        }
    M3 F1i;
    p.J1 = inv3(p.F1, F1i);
    for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += p.dNdx[a][k] * F1i[k][j]; p.dNdy[a][j] = s; }
}

// 6x9 operator T(A,B): rows (11,22,33,12,13,23), columns (11,22,33,12,21,13,31,23,32)
// row (i,i): A(i,k) B(i,l);  row (i,j): A(i,k) B(j,l) + A(j,k) B(i,l)          (continuum_elements.f90:652-669)
void T69(const M3 A, const M3 B, double out[6][9]) {
    static const int RI[6] = {0, 1, 2, 0, 0, 1}, RJ[6] = {0, 1, 2, 1, 2, 2};
    static const int CK[9] = {0, 1, 2, 0, 1, 0, 2, 1, 2}, CL[9] = {0, 1, 2, 1, 0, 2, 0, 2, 1};
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 9; ++c) {
            const int i = RI[r], j = RJ[r], k = CK[c], l = CL[c];
            out[r][c] = (i == j) ? A[i][k] * B[i][l] : A[i][k] * B[j][l] + A[j][k] * B[i][l];
        }
}

// user_codes/src/abaqus_umat_elastic.for:284-305 (3-D branch): ddsdde, stress += ddsdde * dstran
void umat_elastic(const double* props, double stress[6], const double dstran[6], double D[6][6]) {
    const double E = props[0], xnu = props[1];
    std::memset(D, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[i][j] = (i == j) ? 1.0 - xnu : xnu;
    D[3][3] = D[4][4] = D[5][5] = 0.5 * (1.0 - 2.0 * xnu);
    const double f = E / ((1.0 + xnu) * (1.0 - 2.0 * xnu));
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) D[i][j] = D[i][j] * f;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) stress[i] = stress[i] + D[i][j] * dstran[j];
}

}  // namespace

void c3d8_bbar_static(const double* xc, const double* du, const double* ut, const double* matprops, int nprops,
                      double dtime, const double* sv0, double* K, double* R, double* sv1, int nsv) {
    (void)nprops; (void)dtime;
    const int nsp = nsv / 8;                          // state variables per integration point (15 + material states)
    std::memset(K, 0, sizeof(double) * 576);
    std::memset(R, 0, sizeof(double) * 24);
    // ---- volume-averaged quantities (:519-557)
    double J0bar = 0, J1bar = 0, vol = 0, dNbar[8][3];
    static thread_local double Pbar[24][24];
    std::memset(dNbar, 0, sizeof(dNbar));
    std::memset(Pbar, 0, sizeof(Pbar));
    Point p;
    for (int kint = 0; kint < 8; ++kint) {
        point(xc, du, ut, kint, p);
        const double wd = 1.0 * p.det;
        for (int a = 0; a < 8; ++a)
            for (int j = 0; j < 3; ++j) dNbar[a][j] += p.J1 * p.dNdy[a][j] * wd;
        vol += wd;
        J0bar += det3(p.F0) * wd;
        J1bar += p.J1 * wd;
        for (int i = 0; i < 8; ++i)
            for (int r = 0; r < 3; ++r)
                for (int b = 0; b < 8; ++b)
                    for (int k = 0; k < 3; ++k)
                        Pbar[3 * i + r][3 * b + k] += p.J1 * (p.dNdy[i][r] * p.dNdy[b][k] - p.dNdy[i][k] * p.dNdy[b][r]) * wd;
    }
    J0bar /= vol;
    J1bar /= vol;
    for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 3; ++j) dNbar[a][j] /= (J1bar * vol);
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) Pbar[i][j] = Pbar[i][j] / (J1bar * vol) - dNbar[i / 3][i % 3] * dNbar[j / 3][j % 3];

    static const double EYE[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int kint = 0; kint < 8; ++kint) {
        point(xc, du, ut, kint, p);
        const double J0 = det3(p.F0);
        M3 F0, F1, dF, Fmid, Fmi;
        const double s0 = std::pow(J0bar / J0, 1.0 / 3.0), s1 = std::pow(J1bar / p.J1, 1.0 / 3.0);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                F0[i][j] = p.F0[i][j] * s0;
                F1[i][j] = p.F1[i][j] * s1;
                dF[i][j] = F1[i][j] - F0[i][j];
                Fmid[i][j] = 0.5 * (F1[i][j] + F0[i][j]);
            }
        inv3(Fmid, Fmi);
        M3 L, deps, dW, A, B, Bi, dR;
        mul(dF, Fmi, L);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { dW[i][j] = 0.5 * (L[i][j] - L[j][i]); deps[i][j] = 0.5 * (L[i][j] + L[j][i]); }
        for (int i = 0; i < 3; ++i)                                  // Hughes-Winget :591-595
            for (int j = 0; j < 3; ++j) { A[i][j] = EYE[i][j] + 0.5 * dW[i][j]; B[i][j] = EYE[i][j] - 0.5 * dW[i][j]; }
        inv3(B, Bi);
        mul(Bi, A, dR);
        const double* s_in = sv0 + nsp * kint;
        M3 st0 = {{s_in[0], s_in[3], s_in[4]}, {s_in[3], s_in[1], s_in[5]}, {s_in[4], s_in[5], s_in[2]}}, tmp;
        mul(dR, st0, tmp); mulT(tmp, dR, st0);
        double stress[6] = {st0[0][0], st0[1][1], st0[2][2], st0[0][1], st0[0][2], st0[1][2]};
        M3 e0 = {{s_in[6], 0.5 * s_in[9], 0.5 * s_in[10]}, {0.5 * s_in[9], s_in[7], 0.5 * s_in[11]}, {0.5 * s_in[10], 0.5 * s_in[11], s_in[8]}};
        mul(dR, e0, tmp); mulT(tmp, dR, e0);
        const double strain[6] = {e0[0][0], e0[1][1], e0[2][2], e0[0][1], e0[0][2], e0[1][2]};
        const double dstran[6] = {deps[0][0], deps[1][1], deps[2][2], 2.0 * deps[0][1], 2.0 * deps[0][2], 2.0 * deps[1][2]};
        double Dc[6][6];
        umat_elastic(matprops, stress, dstran, Dc);
        // tangent pieces :648-716
        M3 H, Lam, FH, S1, SH, SLam;
        mul(F1, Fmi, FH);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { H[i][j] = FH[j][i]; Lam[i][j] = EYE[i][j] - 0.5 * L[i][j]; }
        double depsdf[6][9], w1[6][9], w2[6][9], Dspin[6][9];
        T69(Lam, H, depsdf);
        M3 sm = {{stress[0], stress[3], stress[4]}, {stress[3], stress[1], stress[5]}, {stress[4], stress[5], stress[2]}};
        std::memcpy(S1, sm, sizeof(M3));
        mul(S1, H, SH);
        mul(S1, Lam, SLam);
        T69(Lam, SH, w1);
        T69(SLam, H, w2);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 9; ++c) Dspin[r][c] = 0.5 * (w1[r][c] - w2[r][c]);
        double* s_out = sv1 + nsp * kint;                            // :718-722
        for (int i = 0; i < 6; ++i) { s_out[i] = stress[i]; s_out[6 + i] = strain[i] + dstran[i]; }
        s_out[12] = s_in[12]; s_out[13] = s_in[13]; s_out[14] = s_in[14];
        for (int i = 15; i < nsp; ++i) s_out[i] = s_in[i];
        // B-bar, B-star :724-755
        double Bb[6][24], Bs[9][24], S[8][3];
        std::memset(Bb, 0, sizeof(Bb));
        std::memset(Bs, 0, sizeof(Bs));
        for (int a = 0; a < 8; ++a) {
            const double* d = p.dNdy[a];
            Bb[0][3 * a] = d[0]; Bb[1][3 * a + 1] = d[1]; Bb[2][3 * a + 2] = d[2];
            Bb[3][3 * a] = d[1]; Bb[3][3 * a + 1] = d[0];
            Bb[4][3 * a] = d[2]; Bb[4][3 * a + 2] = d[0];
            Bb[5][3 * a + 1] = d[2]; Bb[5][3 * a + 2] = d[1];
        }
        for (int a = 0; a < 8; ++a)
            for (int k = 0; k < 3; ++k) { double s = 0; for (int r = 0; r < 6; ++r) s += Bb[r][3 * a + k] * stress[r]; S[a][k] = s; }
        for (int i = 0; i < 3; ++i)
            for (int a = 0; a < 8; ++a)
                for (int k = 0; k < 3; ++k) Bb[i][3 * a + k] += (dNbar[a][k] - p.dNdy[a][k]) / 3.0;
        for (int i = 0; i < 3; ++i) std::memcpy(Bs[i], Bb[i], sizeof(double) * 24);
        for (int a = 0; a < 8; ++a) {
            const double* d = p.dNdy[a];
            Bs[3][3 * a] = d[1]; Bs[4][3 * a + 1] = d[0]; Bs[5][3 * a] = d[2];
            Bs[6][3 * a + 2] = d[0]; Bs[7][3 * a + 1] = d[2]; Bs[8][3 * a + 2] = d[1];
        }
        const double press = (stress[0] + stress[1] + stress[2]) / 3.0;
        const double wdj = 1.0 * p.det * J1bar;
        for (int i = 0; i < 24; ++i) {
            double s = 0; for (int r = 0; r < 6; ++r) s += Bb[r][i] * stress[r];
            R[i] = R[i] - s * wdj;
        }
        // G = (Dcorot * depsdf + Dspin) * Bstar  (6x24); K += Bbar^T G
        double M[6][9], G[6][24];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 9; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += Dc[r][k] * depsdf[k][c]; M[r][c] = s + Dspin[r][c]; }
        for (int r = 0; r < 6; ++r)
            for (int j = 0; j < 24; ++j) { double s = 0; for (int c = 0; c < 9; ++c) s += M[r][c] * Bs[c][j]; G[r][j] = s; }
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                double s = 0; for (int r = 0; r < 6; ++r) s += Bb[r][i] * G[r][j];
                const int a = i / 3, ri = i % 3, b = j / 3, kj = j % 3;
                const double pm_sm = p.dNdy[a][kj] * S[b][ri];               // (Pmat o Smat^T)
                const double pm_pm = p.dNdy[a][kj] * p.dNdy[b][ri];          // (Pmat o Pmat^T)
                K[i + 24 * j] += (s - pm_sm + press * (Pbar[i][j] + pm_pm)) * wdj;
            }
    }
}

// Gibbs-Poole-Stockmeyer renumbering (src/bandwidth_minimizer.f90) is not restated: the skyline uses the identity
// node order, which changes the profile statistics and LU conditioning lines of the golden log but not the solution.
void gps_node_order(const Model& m, std::vector<int>& node_order) {
    node_order.resize(m.n_nodes);
    for (int i = 0; i < m.n_nodes; ++i) node_order[i] = i;
}

}  // namespace oracle
