// ORACLE — test infrastructure only (see or_model.h).  Element routines.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "or_model.h"

namespace oracle {

// Node sign table of the 8-node brick, ABAQUS C3D8 order
// (user_codes/modules/Element_Utilities.f90:908-940 = abaqus_uel_linelastic_3d.for:424-456).
static const double SGN[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                 {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};

// Gauss abscissa: the reference assigns the *single precision* literal 0.5773502692 to a double
// (Element_Utilities.f90:637-638, abaqus_uel_linelastic_3d.for:309-310) => 0.57735025882720947.
static const double GP = (double)0.5773502692f;

// 2x2x2 rule, point n = 4(k-1)+2(j-1)+i with i fastest over xi_1 (Element_Utilities.f90:639-648), w = 1.
static void gauss_point(int n, double xi[3]) {
    int i = n & 1, j = (n >> 1) & 1, k = (n >> 2) & 1;
    xi[0] = i ? GP : -GP;
    xi[1] = j ? GP : -GP;
    xi[2] = k ? GP : -GP;
}

// dN/dxi for hex8 (Element_Utilities.f90:916-939).  (1 -/+ xi) products then /8, sign applied last:
// bit-identical to the reference's "-(1.-xi(2))*(1.-xi(3))/8." forms.
static void shape_hex8(const double xi[3], double N[8], double dNdxi[8][3]) {
    for (int a = 0; a < 8; ++a) {
        double f1 = 1.0 + SGN[a][0] * xi[0];
        double f2 = 1.0 + SGN[a][1] * xi[1];
        double f3 = 1.0 + SGN[a][2] * xi[2];
        N[a] = f1 * f2 * f3 / 8.0;
        dNdxi[a][0] = SGN[a][0] * (f2 * f3 / 8.0);
        dNdxi[a][1] = SGN[a][1] * (f1 * f3 / 8.0);
        dNdxi[a][2] = SGN[a][2] * (f1 * f2 / 8.0);
    }
}

// Cofactor inverse + determinant (Element_Utilities.f90:99-123 / abq_UEL_invert3d :573-608).
// A[i][j] row i, column j.  Returns 0 on a zero determinant (the reference stops).
static int invert3(const double A[3][3], double Ai[3][3], double* det) {
    double d = A[0][0] * A[1][1] * A[2][2] - A[0][0] * A[1][2] * A[2][1] - A[0][1] * A[1][0] * A[2][2] +
               A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] - A[0][2] * A[1][1] * A[2][0];
    *det = d;
    if (d == 0.0) return 0;
    double C[3][3];
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
    return 1;
}

// Geometry at one Gauss point: dNdx(8,3) = dNdxi * dxidx with dxdxi = x(3x8)*dNdxi(8x3)
// (el_linelast_3Dbasic.f90:109-112).
static int geometry(const double* coords24, int kint, double dNdx[8][3], double* det, double N[8]) {
    double xi[3], dNdxi[8][3], J[3][3], Ji[3][3];
    gauss_point(kint, xi);
    shape_hex8(xi, N, dNdxi);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int a = 0; a < 8; ++a) s += coords24[3 * a + i] * dNdxi[a][j];
            J[i][j] = s;
        }
    if (!invert3(J, Ji, det)) return 0;
    for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += dNdxi[a][k] * Ji[k][j];
            dNdx[a][j] = s;
        }
    return 1;
}

// B(6x24): rows e11,e22,e33,2e12,2e13,2e23 (el_linelast_3Dbasic.f90:113-122)
static void bmatrix(const double dNdx[8][3], double B[6][24]) {
    std::memset(B, 0, sizeof(double) * 6 * 24);
    for (int a = 0; a < 8; ++a) {
        B[0][3 * a + 0] = dNdx[a][0];
        B[1][3 * a + 1] = dNdx[a][1];
        B[2][3 * a + 2] = dNdx[a][2];
        B[3][3 * a + 0] = dNdx[a][1];
        B[3][3 * a + 1] = dNdx[a][0];
        B[4][3 * a + 0] = dNdx[a][2];
        B[4][3 * a + 2] = dNdx[a][0];
        B[5][3 * a + 1] = dNdx[a][2];
        B[5][3 * a + 2] = dNdx[a][1];
    }
}

static void dmatrix(const double* props, double D[6][6]) {
    // el_linelast_3Dbasic.f90:93-105
    double E = props[0], xnu = props[1];
    double d44 = 0.5 * E / (1 + xnu);
    double d11 = (1.0 - xnu) * E / ((1 + xnu) * (1 - 2.0 * xnu));
    double d12 = xnu * E / ((1 + xnu) * (1 - 2.0 * xnu));
    std::memset(D, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[i][j] = d12;
    D[0][0] = D[1][1] = D[2][2] = d11;
    D[3][3] = D[4][4] = D[5][5] = d44;
}

static void accumulate_KR(const double B[6][24], const double D[6][6], const double stress[6], double wdet_w,
                          double det, double* K, double* R) {
    // R -= matmul(B^T, stress)*w*det ; K += matmul(B^T, matmul(D,B))*w*det  (el_linelast_3Dbasic.f90:128-131)
    double DB[6][24];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 24; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += D[i][k] * B[k][j];
            DB[i][j] = s;
        }
    for (int i = 0; i < 24; ++i) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += B[k][i] * stress[k];
        R[i] = R[i] - s * wdet_w * det;
    }
    for (int j = 0; j < 24; ++j)
        for (int i = 0; i < 24; ++i) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += B[k][i] * DB[k][j];
            K[i + 24 * j] = K[i + 24 * j] + s * wdet_w * det;
        }
}

// Native element id 1001 (user_codes/src/el_linelast_3Dbasic.f90:79-133), hex8 branch only.
void el_linelast_3dbasic(const double* coords24, const double* du24, const double* utot24,
                         const double* props, double* K, double* R) {
    double D[6][6], B[6][24], dNdx[8][3], N[8], det;
    std::memset(K, 0, sizeof(double) * 576);
    std::memset(R, 0, sizeof(double) * 24);
    dmatrix(props, D);
    for (int kint = 0; kint < 8; ++kint) {
        if (!geometry(coords24, kint, dNdx, &det, N)) {
            std::fprintf(stderr, "oracle: zero Jacobian determinant (Element_Utilities.f90:108-112)\n");
            return;
        }
        bmatrix(dNdx, B);
        double strain[6], dstrain[6], stress[6];
        for (int i = 0; i < 6; ++i) {
            double s = 0.0, ds = 0.0;
            for (int k = 0; k < 24; ++k) s += B[i][k] * utot24[k];
            for (int k = 0; k < 24; ++k) ds += B[i][k] * du24[k];
            strain[i] = s;
            dstrain[i] = ds;
        }
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += D[i][k] * (strain[k] + dstrain[k]);
            stress[i] = s;
        }
        accumulate_KR(B, D, stress, 1.0, det, K, R);
    }
}

// Face node table for hex8, 1-based local nodes (Element_Utilities.f90:1102-1109 =
// abq_facenodes_3D abaqus_uel_linelastic_3d.for:645-651).
void facenodes_hex8(int face, int list[4]) {
    static const int T[6][4] = {{1, 2, 3, 4}, {5, 8, 7, 6}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 8, 4}, {4, 8, 5, 1}};
    for (int i = 0; i < 4; ++i) list[i] = T[face - 1][i];
}

// 2x2 face rule with the *proper double* abscissa (abaqus_uel_2D.for:190 = Element_Utilities.f90:486)
static void face_gauss(int k, double xi[2]) {
    const double cn = 0.5773502691896260;
    static const int sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
    xi[0] = sx[k] * cn;
    xi[1] = sy[k] * cn;
}
// 4-node quad shape functions (abaqus_uel_2D.for:311-334)
static void shape_quad4(const double xi[2], double f[4], double df[4][2]) {
    double g1 = 0.5 * (1.0 - xi[0]), g2 = 0.5 * (1.0 + xi[0]);
    double h1 = 0.5 * (1.0 - xi[1]), h2 = 0.5 * (1.0 + xi[1]);
    f[0] = g1 * h1; f[1] = g2 * h1; f[2] = g2 * h2; f[3] = g1 * h2;
    const double dg1 = -0.5, dg2 = 0.5, dh1 = -0.5, dh2 = 0.5;
    df[0][0] = dg1 * h1; df[1][0] = dg2 * h1; df[2][0] = dg2 * h2; df[3][0] = dg1 * h2;
    df[0][1] = g1 * dh1; df[1][1] = g2 * dh1; df[2][1] = g2 * dh2; df[3][1] = g1 * dh2;
}
static void face_normal(const double* fc12, const double df[4][2], double norm[3]) {
    double dx[3][2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j) {
            double s = 0.0;
            for (int a = 0; a < 4; ++a) s += fc12[3 * a + i] * df[a][j];
            dx[i][j] = s;
        }
    // sign of norm(2) as written in the reference (abaqus_uel_linelastic_3d.for:228-230, Q2)
    norm[0] = (dx[1][0] * dx[2][1]) - (dx[1][1] * dx[2][0]);
    norm[1] = (dx[0][0] * dx[2][1]) - (dx[0][1] * dx[2][0]);
    norm[2] = (dx[0][0] * dx[1][1]) - (dx[0][1] * dx[1][0]);
}

// ABAQUS-format UEL "U1" (user_codes/src/abaqus_uel_linelastic_3d.for:129-238), hex8 branch.
// U24 = u + du as passed at solver_direct.f90:234.
void uel_linelastic_3d(const double* coords24, const double* U24, const double* props, int nsvars,
                       int ndload, const int* jdltyp, const double* adlmag, double* K, double* R,
                       double* svars, double* energy8, double* pnewdt) {
    double D[6][6], B[6][24], dNdx[8][3], N[8], det;
    std::memset(K, 0, sizeof(double) * 576);
    std::memset(R, 0, sizeof(double) * 24);
    if (energy8) std::memset(energy8, 0, sizeof(double) * 8);
    dmatrix(props, D);
    for (int kint = 0; kint < 8; ++kint) {
        if (!geometry(coords24, kint, dNdx, &det, N)) {
            std::fprintf(stderr, "oracle: zero Jacobian determinant (abaqus_uel_linelastic_3d.for:589-593)\n");
            return;
        }
        bmatrix(dNdx, B);
        double strain[6], stress[6];
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int k = 0; k < 24; ++k) s += B[i][k] * U24[k];
            strain[i] = s;
        }
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += D[i][k] * strain[k];
            stress[i] = s;
        }
        accumulate_KR(B, D, stress, 1.0, det, K, R);
        if (energy8) {
            double e = 0.0;
            for (int i = 0; i < 6; ++i) e += stress[i] * strain[i];
            energy8[1] = energy8[1] + 0.5 * e * 1.0 * det;                     // :190-191
        }
        if (svars && nsvars >= 48)
            for (int i = 0; i < 6; ++i) svars[6 * kint + i] = stress[i];       // :193-195
    }
    if (pnewdt) *pnewdt = 1.0;                                                  // :199
    // Distributed (pressure) loads :207-238.  The update at :234-235 multiplies N2(1:nfacenodes) by
    // norm(1:3); scalarised over the 3-element left-hand side every face node receives
    // (N2(1) n1, N2(2) n2, N2(3) n3).  Reproduced literally.
    for (int j = 0; j < ndload; ++j) {
        int list[4];
        facenodes_hex8(jdltyp[j] < 0 ? -jdltyp[j] : jdltyp[j], list);
        double fc[12];
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 3; ++c) fc[3 * i + c] = coords24[3 * (list[i] - 1) + c];
        for (int kint = 0; kint < 4; ++kint) {
            double xi2[2], N2[4], dN2[4][2], norm[3];
            face_gauss(kint, xi2);
            shape_quad4(xi2, N2, dN2);
            face_normal(fc, dN2, norm);
            for (int i = 0; i < 4; ++i) {
                int ip = 3 * (list[i] - 1);
                for (int c = 0; c < 3; ++c) R[ip + c] = R[ip + c] - N2[c] * adlmag[j] * norm[c] * 1.0;
            }
        }
    }
}

// Surface traction on a native-element face (src/traction_boundarycondition.f90:66-94), 3-D, 4-node face.
// flag 1/2: traction vector given; flag 3: normal pressure traction(1).
void traction_boundarycondition_static_3d(int flag, const double* fc12, const double* traction, int ntract,
                                          double* R12) {
    std::memset(R12, 0, sizeof(double) * 12);
    for (int kint = 0; kint < 4; ++kint) {
        double xi2[2], N2[4], dN2[4][2], norm[3];
        face_gauss(kint, xi2);
        shape_quad4(xi2, N2, dN2);
        face_normal(fc12, dN2, norm);
        double det = std::sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
        if (flag < 3) {
            for (int k = 0; k < 3; ++k) {
                double t = k < ntract ? traction[k] : 0.0;
                for (int a = 0; a < 4; ++a) R12[3 * a + k] = R12[3 * a + k] + N2[a] * t * 1.0 * det;
            }
        } else {
            for (int a = 0; a < 4; ++a)
                for (int k = 0; k < 3; ++k) R12[3 * a + k] = R12[3 * a + k] - N2[a] * traction[0] * norm[k] * 1.0;
        }
    }
}

// traction_boundarycondition_dynamic (src/traction_boundarycondition.f90:102-198), 3-D 4-node face.  flag 1/2 as the static
// routine; flag 3 (normal) as written there: + traction(1) norm(1), - traction(2) norm(2), - traction(3) norm(3), with the
// caller's traction vector (explicit_dynamic_step.f90:905-920 passes (dloadvalue, 0, 0): only the first term survives).
void traction_boundarycondition_dynamic_3d(int flag, const double* fc12, const double* traction3, int ntract, double* R12) {
    std::memset(R12, 0, sizeof(double) * 12);
    for (int kint = 0; kint < 4; ++kint) {
        double xi2[2], N2[4], dN2[4][2], norm[3];
        face_gauss(kint, xi2);
        shape_quad4(xi2, N2, dN2);
        face_normal(fc12, dN2, norm);
        double det = std::sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
        if (flag < 3) {
            for (int k = 0; k < 3; ++k) {
                double t = k < ntract ? traction3[k] : 0.0;
                for (int a = 0; a < 4; ++a) R12[3 * a + k] = R12[3 * a + k] + N2[a] * t * 1.0 * det;
            }
        } else {
            for (int a = 0; a < 4; ++a) {
                R12[3 * a] = R12[3 * a] + N2[a] * traction3[0] * norm[0] * 1.0;
                R12[3 * a + 1] = R12[3 * a + 1] - N2[a] * traction3[1] * norm[1] * 1.0;
                R12[3 * a + 2] = R12[3 * a + 2] - N2[a] * traction3[2] * norm[2] * 1.0;
            }
        }
    }
}

// VUEL distributed load (user_codes/src/abaqus_vuel.for:274-320): RHS(ipoin:ipoin+2) -= N2(i) adlmag norm(1:3) w
void vuel_pressure_3d(const double* coords24, int face, double adlmag, double* R24) {
    std::memset(R24, 0, sizeof(double) * 24);
    int list[4];
    facenodes_hex8(face < 0 ? -face : face, list);
    double fc[12];
    for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 3; ++c) fc[3 * i + c] = coords24[3 * (list[i] - 1) + c];
    for (int kint = 0; kint < 4; ++kint) {
        double xi2[2], N2[4], dN2[4][2], norm[3];
        face_gauss(kint, xi2);
        shape_quad4(xi2, N2, dN2);
        face_normal(fc, dN2, norm);
        for (int i = 0; i < 4; ++i) {
            int ip = 3 * (list[i] - 1);
            for (int c = 0; c < 3; ++c) R24[ip + c] = R24[ip + c] - N2[i] * adlmag * norm[c] * 1.0;
        }
    }
}

// History table interpolation (modules/Boundaryconditions.f90:161-221)
double interpolate_history_table(const double* t, const double* v, int n, double time) {
    if (n == 1) return v[0];
    if (time <= t[0]) return v[0];
    if (time >= t[n - 1]) return v[n - 1];
    int klo = 0, khi = n - 1;
    while (khi - klo > 1) {
        int k = (khi + klo + 2) / 2 - 1;   // same midpoint as the 1-based (khi+klo)/2
        if (t[k] > time) khi = k; else klo = k;
    }
    if (t[khi] == t[klo]) return 0.5 * (v[khi] + v[klo]);
    return ((t[khi] - time) * v[klo] + (time - t[klo]) * v[khi]) / (t[khi] - t[klo]);
}

// ---------------------------------------------------------------------------------------------
// Compressible neo-Hookean hex8 (total Lagrangian), W = mu/2 (Ibar1 - 3) + K/2 (J - 1)^2,
// Ibar1 = J^(-2/3) tr(F^T F).  props = (mu, K [, ignored...]) as in
// input_files/Abaqus_uel_hyperelastic.in:44 (G11..G44 ignored).  NOT in the reference (no
// hyperelastic code is shipped): parity unpinned by the reference; pinned by the FD stiffness check
// (checkstiffness.f90:481-532 criterion) and patch tests.  Same Gauss rule/shape functions as above.
//   P = dW/dF = mu J^(-2/3) (F - I1/3 F^-T) + K (J-1) J F^-T
//   R_ai = - sum_gp P_iJ dN_a/dX_J w|J0| ;  K_ai,bk = sum_gp dN_a/dX_J A_iJkL dN_b/dX_L w|J0|
void el_neohooke_3d(const double* coords24, const double* U24, const double* props, double* K,
                    double* R, double* svars, int* fail) {
    const double mu = props[0], Kb = props[1];
    std::memset(K, 0, sizeof(double) * 576);
    std::memset(R, 0, sizeof(double) * 24);
    *fail = 0;
    for (int kint = 0; kint < 8; ++kint) {
        double dNdX[8][3], N[8], det0;
        if (!geometry(coords24, kint, dNdX, &det0, N)) { *fail = 1; return; }
        double F[3][3], Fi[3][3], J;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = (i == j) ? 1.0 : 0.0;
                for (int a = 0; a < 8; ++a) s += U24[3 * a + i] * dNdX[a][j];
                F[i][j] = s;
            }
        if (!invert3(F, Fi, &J) || !(J > 0.0)) { *fail = 1; return; }
        double I1 = 0.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) I1 += F[i][j] * F[i][j];
        const double Jm23 = std::pow(J, -2.0 / 3.0);
        double P[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double FiT = Fi[j][i];
                P[i][j] = mu * Jm23 * (F[i][j] - I1 / 3.0 * FiT) + Kb * (J - 1.0) * J * FiT;
            }
        if (svars) {
            // store Cauchy stress sigma = P F^T / J (6 components s11,s22,s33,s12,s13,s23)
            double sg[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double s = 0.0;
                    for (int k = 0; k < 3; ++k) s += P[i][k] * F[j][k];
                    sg[i][j] = s / J;
                }
            double* sv = svars + 6 * kint;
            sv[0] = sg[0][0]; sv[1] = sg[1][1]; sv[2] = sg[2][2]; sv[3] = sg[0][1]; sv[4] = sg[0][2]; sv[5] = sg[1][2];
        }
        const double w = 1.0 * det0;
        for (int a = 0; a < 8; ++a)
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
                for (int j = 0; j < 3; ++j) s += P[i][j] * dNdX[a][j];
                R[3 * a + i] -= s * w;
            }
        // material tangent A_iJkL = dP_iJ/dF_kL
        //  = mu Jm23 [ d_ik d_JL - (2/3) F_iJ Fi_Lk - (2/3) Fi_Ji F_kL + (2/9) I1 Fi_Ji Fi_Lk + (I1/3) Fi_Jk Fi_Li ]
        //  + K [ (2J-1) J Fi_Ji Fi_Lk - (J-1) J Fi_Jk Fi_Li ]
        const double c1 = mu * Jm23, cJ = Kb * (2.0 * J - 1.0) * J, cK = Kb * (J - 1.0) * J;
        for (int a = 0; a < 8; ++a)
            for (int b = 0; b < 8; ++b)
                for (int i = 0; i < 3; ++i)
                    for (int k = 0; k < 3; ++k) {
                        double s = 0.0;
                        for (int Jx = 0; Jx < 3; ++Jx)
                            for (int L = 0; L < 3; ++L) {
                                double A = c1 * (((i == k && Jx == L) ? 1.0 : 0.0) -
                                                 (2.0 / 3.0) * F[i][Jx] * Fi[L][k] -
                                                 (2.0 / 3.0) * Fi[Jx][i] * F[k][L] +
                                                 (2.0 / 9.0) * I1 * Fi[Jx][i] * Fi[L][k] +
                                                 (I1 / 3.0) * Fi[Jx][k] * Fi[L][i]) +
                                           cJ * Fi[Jx][i] * Fi[L][k] - cK * Fi[Jx][k] * Fi[L][i];
                                s += dNdX[a][Jx] * A * dNdX[b][L];
                            }
                        K[(3 * a + i) + 24 * (3 * b + k)] += s * w;
                    }
    }
}

// Element dispatch + gather, mirroring solver_direct.f90:104-258 (gather :104-121, UEL :174-247,
// native :250-256).  UEL distributed loads are generated as in generate_abaqus_dloads :798-912.
// ---- integration rules and shape functions of the other 3-D element shapes, AS THE REFERENCE WRITES THEM -----------------
// initialize_integration_points (Element_Utilities.f90:586-689): literals without a D exponent are single precision
// there (0.58541020, 0.13819660, 0.7745966692, ...) and are reproduced as (double)(float) values.
int integration_points_3d(int n_nodes, double xi[][3], double* w) {
    if (n_nodes == 4) {                                   // 1 point (el_linelast_3Dbasic.f90:82)
        xi[0][0] = xi[0][1] = xi[0][2] = 0.25;
        w[0] = 1.0 / 6.0;
        return 1;
    }
    if (n_nodes == 10) {                                  // 4 points (:83)
        const double a = (double)0.58541020f, b = (double)0.13819660f;
        const double p[4][3] = {{a, b, b}, {b, a, b}, {b, b, a}, {b, b, b}};
        for (int k = 0; k < 4; ++k) { for (int j = 0; j < 3; ++j) xi[k][j] = p[k][j]; w[k] = 1.0 / 24.0; }
        return 4;
    }
    if (n_nodes == 8) {
        const double x1[2] = {(double)-0.5773502692f, (double)0.5773502692f};
        for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) {
            const int n = 4 * k + 2 * j + i;
            xi[n][0] = x1[i]; xi[n][1] = x1[j]; xi[n][2] = x1[k]; w[n] = 1.0;
        }
        return 8;
    }
    if (n_nodes == 20) {                                  // 27 points (:85)
        const double x1[3] = {(double)-0.7745966692f, 0.0, (double)0.7745966692f};
        const double w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
        for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) {
            const int n = 9 * k + 3 * j + i;
            xi[n][0] = x1[i]; xi[n][1] = x1[j]; xi[n][2] = x1[k]; w[n] = w1[i] * w1[j] * w1[k];
        }
        return 27;
    }
    return 0;
}

// calculate_shapefunctions, 3-D branch (Element_Utilities.f90:863-1025), statement by statement.  The reference's table
// has three entries that are not the derivatives of its own shape functions — tet10 df(10,1) = -4 xi3 xi4 (:902) and
// hex20 df(11,1), df(11,2) (copies of node 9's, :991-992): they are reproduced, because "the same results as the
// reference on the same input" is the bar; tests/test_oracle_cpu.py pins this table against the Fortran text itself.
void shapefunctions_3d(int n_nodes, const double* xi, double* f, double df[][3]) {
    for (int a = 0; a < n_nodes; ++a) { f[a] = 0.0; df[a][0] = df[a][1] = df[a][2] = 0.0; }
    const double x = xi[0], y = xi[1], z = xi[2];
    if (n_nodes == 4) {
        f[0] = x; f[1] = y; f[2] = z; f[3] = 1. - x - y - z;
        df[0][0] = 1.; df[1][1] = 1.; df[2][2] = 1.;
        df[3][0] = -1.; df[3][1] = -1.; df[3][2] = -1.;
    } else if (n_nodes == 10) {
        const double x4 = 1. - x - y - z;
        f[0] = (2. * x - 1.) * x; f[1] = (2. * y - 1.) * y; f[2] = (2. * z - 1.) * z; f[3] = (2. * x4 - 1.) * x4;
        f[4] = 4. * x * y; f[5] = 4. * y * z; f[6] = 4. * z * x; f[7] = 4. * x * x4; f[8] = 4. * y * x4; f[9] = 4. * z * x4;
        df[0][0] = (4. * x - 1.); df[1][1] = (4. * y - 1.); df[2][2] = (4. * z - 1.);
        df[3][0] = -(4. * x4 - 1.); df[3][1] = -(4. * x4 - 1.); df[3][2] = -(4. * x4 - 1.);
        df[4][0] = 4. * y; df[4][1] = 4. * x;
        df[5][1] = 4. * z; df[5][2] = 4. * y;
        df[6][0] = 4. * z; df[6][2] = 4. * x;
        df[7][0] = 4. * (x4 - x); df[7][1] = -4. * x; df[7][2] = -4. * x;
        df[8][0] = -4. * y; df[8][1] = 4. * (x4 - y); df[8][2] = -4. * y;
        df[9][0] = -4. * z * x4;                           // sic (:902)
        df[9][1] = -4. * z; df[9][2] = 4. * (x4 - z);
    } else if (n_nodes == 8) {
        const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        for (int a = 0; a < 8; ++a) {
            const double fx = 1. + sx[a] * x, fy = 1. + sy[a] * y, fz = 1. + sz[a] * z;
            f[a] = fx * fy * fz / 8.;
            df[a][0] = sx[a] * (fy * fz / 8.); df[a][1] = sy[a] * (fx * fz / 8.); df[a][2] = sz[a] * (fx * fy / 8.);
        }
    } else if (n_nodes == 20) {
        const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        for (int a = 0; a < 8; ++a) {                      // corner nodes (:926-933, :946-973)
            const double fx = 1. + sx[a] * x, fy = 1. + sy[a] * y, fz = 1. + sz[a] * z;
            const double g = sx[a] * x + sy[a] * y + sz[a] * z - 2.;
            f[a] = fx * fy * fz * g / 8.;
            df[a][0] = (sx[a] * (fy * fz * g) + sx[a] * (fx * fy * fz)) / 8.;
            df[a][1] = (sy[a] * (fx * fz * g) + sy[a] * (fx * fy * fz)) / 8.;
            df[a][2] = (sz[a] * (fx * fy * g) + sz[a] * (fx * fy * fz)) / 8.;
        }
        const double xm = 1. - x, xp = 1. + x, ym = 1. - y, yp = 1. + y, zm = 1. - z, zp = 1. + z;
        const double x2 = 1. - x * x, y2 = 1. - y * y, z2 = 1. - z * z;
        f[8] = x2 * ym * zm / 4.;  f[9] = xp * y2 * zm / 4.;  f[10] = x2 * yp * zm / 4.; f[11] = xm * y2 * zm / 4.;
        f[12] = x2 * ym * zp / 4.; f[13] = xp * y2 * zp / 4.; f[14] = x2 * yp * zp / 4.; f[15] = xm * y2 * zp / 4.;
        f[16] = xm * ym * z2 / 4.; f[17] = xp * ym * z2 / 4.; f[18] = xp * yp * z2 / 4.; f[19] = xm * yp * z2 / 4.;
        df[8][0] = -2. * x * ym * zm / 4.;  df[8][1] = -x2 * zm / 4.;            df[8][2] = -x2 * ym / 4.;
        df[9][0] = y2 * zm / 4.;            df[9][1] = -2. * y * xp * zm / 4.;   df[9][2] = -y2 * xp / 4.;
        df[10][0] = -2. * x * ym * zm / 4.; df[10][1] = -x2 * zm / 4.;           df[10][2] = -x2 * ym / 4.;   // sic (:991-993)
        df[11][0] = -y2 * zm / 4.;          df[11][1] = -2. * y * xm * zm / 4.;  df[11][2] = -y2 * xm / 4.;
        df[12][0] = -2. * x * ym * zp / 4.; df[12][1] = -x2 * zp / 4.;           df[12][2] = x2 * ym / 4.;
        df[13][0] = y2 * zp / 4.;           df[13][1] = -2. * y * xp * zp / 4.;  df[13][2] = y2 * xp / 4.;
        df[14][0] = 2. * x * yp * zp / 4.;  df[14][1] = x2 * zp / 4.;            df[14][2] = x2 * yp / 4.;    // sic (:1003): + 2 xi1
        df[15][0] = -y2 * zp / 4.;          df[15][1] = -2. * y * xm * zp / 4.;  df[15][2] = y2 * xm / 4.;
        df[16][0] = -ym * z2 / 4.;          df[16][1] = -xm * z2 / 4.;           df[16][2] = -z * xm * ym / 2.;
        df[17][0] = ym * z2 / 4.;           df[17][1] = -xp * z2 / 4.;           df[17][2] = -z * xp * ym / 2.;
        df[18][0] = yp * z2 / 4.;           df[18][1] = xp * z2 / 4.;            df[18][2] = -z * xp * yp / 2.;
        df[19][0] = -yp * z2 / 4.;          df[19][1] = xm * z2 / 4.;            df[19][2] = -z * xm * yp / 2.;
    }
}

// el_linelast_3dbasic for any of the four shapes (el_linelast_3Dbasic.f90:79-133): B^T D B and B^T sigma summed over the
// integration points in the reference's order.
int el_linelast_3d_general(int nn, const double* coords, const double* du, const double* utot, const double* props,
                           double* K, double* R) {
    const int nd = 3 * nn;
    double xi[27][3], w[27];
    const int npts = integration_points_3d(nn, xi, w);
    if (npts == 0) return 1;
    for (int q = 0; q < nd * nd; ++q) K[q] = 0.0;
    for (int q = 0; q < nd; ++q) R[q] = 0.0;
    const double E = props[0], xnu = props[1];
    const double d44 = 0.5 * E / (1 + xnu), d11 = (1. - xnu) * E / ((1 + xnu) * (1 - 2. * xnu)), d12 = xnu * E / ((1 + xnu) * (1 - 2. * xnu));
    double D[6][6] = {{0}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) D[i][j] = d12;
    D[0][0] = D[1][1] = D[2][2] = d11;
    D[3][3] = D[4][4] = D[5][5] = d44;
    int fail = 0;
    for (int k = 0; k < npts; ++k) {
        double f[20], df[20][3], dxdxi[3][3], dxidx[3][3], dNdx[20][3], B[6][60];
        shapefunctions_3d(nn, xi[k], f, df);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double t = 0.0;
                for (int a = 0; a < nn; ++a) t += coords[3 * a + i] * df[a][j];
                dxdxi[i][j] = t;
            }
        const double (*A)[3] = dxdxi;                                   // invert_small, Element_Utilities.f90:99-123
        const double det = A[0][0] * A[1][1] * A[2][2] - A[0][0] * A[1][2] * A[2][1] - A[0][1] * A[1][0] * A[2][2] +
                           A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] - A[0][2] * A[1][1] * A[2][0];
        if (det == 0.0) fail = 1;
        const double r = 1.0 / det;
        dxidx[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) * r;  dxidx[0][1] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]) * r;
        dxidx[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * r;  dxidx[1][0] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]) * r;
        dxidx[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * r;  dxidx[1][2] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]) * r;
        dxidx[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) * r;  dxidx[2][1] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]) * r;
        dxidx[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * r;
        for (int a = 0; a < nn; ++a)
            for (int j = 0; j < 3; ++j) dNdx[a][j] = df[a][0] * dxidx[0][j] + df[a][1] * dxidx[1][j] + df[a][2] * dxidx[2][j];
        for (int r6 = 0; r6 < 6; ++r6) for (int q = 0; q < nd; ++q) B[r6][q] = 0.0;
        for (int a = 0; a < nn; ++a) {
            B[0][3 * a] = dNdx[a][0]; B[1][3 * a + 1] = dNdx[a][1]; B[2][3 * a + 2] = dNdx[a][2];
            B[3][3 * a] = dNdx[a][1]; B[3][3 * a + 1] = dNdx[a][0];
            B[4][3 * a] = dNdx[a][2]; B[4][3 * a + 2] = dNdx[a][0];
            B[5][3 * a + 1] = dNdx[a][2]; B[5][3 * a + 2] = dNdx[a][1];
        }
        double strain[6], stress[6], DB[6][60];
        for (int r6 = 0; r6 < 6; ++r6) {
            double s0 = 0.0, s1 = 0.0;
            for (int q = 0; q < nd; ++q) { s0 += B[r6][q] * utot[q]; s1 += B[r6][q] * du[q]; }
            strain[r6] = s0 + s1;
        }
        for (int r6 = 0; r6 < 6; ++r6) { double t = 0.0; for (int c = 0; c < 6; ++c) t += D[r6][c] * strain[c]; stress[r6] = t; }
        const double wd = w[k] * det;
        for (int q = 0; q < nd; ++q) {
            double t = 0.0;
            for (int r6 = 0; r6 < 6; ++r6) t += B[r6][q] * stress[r6];
            R[q] = R[q] - t * wd;
        }
        for (int r6 = 0; r6 < 6; ++r6)
            for (int q = 0; q < nd; ++q) { double t = 0.0; for (int c = 0; c < 6; ++c) t += D[r6][c] * B[c][q]; DB[r6][q] = t; }
        for (int i = 0; i < nd; ++i)
            for (int j = 0; j < nd; ++j) {
                double t = 0.0;
                for (int r6 = 0; r6 < 6; ++r6) t += B[r6][i] * DB[r6][j];
                K[i + nd * j] += t * wd;
            }
    }
    return fail;
}

int element_static(const Model& m, int lmn, const double* dof_total, const double* dof_increment,
                   double time, double dtime, double* K, double* R, double* svars) {
    if (m.nodes_per_element != 8) {                      // tet4 / tet10 / hex20 meshes: element 1001 only
        const int npe = m.nodes_per_element;
        double gx[60], gdu[60], gut[60];
        for (int j = 0; j < npe; ++j) {
            int n = m.connectivity[(size_t)npe * lmn + j] - 1;
            for (int k = 0; k < 3; ++k) {
                gx[3 * j + k] = m.coords[3 * n + k];
                gdu[3 * j + k] = dof_increment[3 * n + k];
                gut[3 * j + k] = dof_total[3 * n + k];
            }
        }
        if (m.element_flag[lmn] != ID_LINELAST) return 1;
        return el_linelast_3d_general(npe, gx, gdu, gut, m.element_properties.data() + m.element_prop_index[lmn], K, R);
    }
    double xc[24], du[24], ut[24];
    for (int j = 0; j < 8; ++j) {
        int n = m.connectivity[8 * lmn + j] - 1;
        for (int k = 0; k < 3; ++k) {
            xc[3 * j + k] = m.coords[3 * n + k];
            du[3 * j + k] = dof_increment[3 * n + k];
            ut[3 * j + k] = dof_total[3 * n + k];
        }
    }
    const int flag = m.element_flag[lmn];
    const double* props = m.element_properties.data() + m.element_prop_index[lmn];
    if (flag == FLAG_C3D8) {
        int mat = m.element_material[lmn] - 1;
        if (!svars) { std::fprintf(stderr, "oracle: C3D8 element needs state variable storage\n"); return 1; }
        // svars holds the state at the start of the increment on entry and is updated in place (each integration
        // point reads its initial state before writing the updated one)
        c3d8_bbar_static(xc, du, ut, m.material_properties.data() + m.material_prop_index[mat],
                         m.material_n_props[mat], dtime, svars, K, R, svars, m.element_n_states[lmn]);
        return 0;
    }
    if (flag > FLAG_UEL_BASE) {
        double U[24];
        for (int k = 0; k < 24; ++k) U[k] = du[k] + ut[k];          // solver_direct.f90:234
        // distributed loads on this element, in load order (generate_abaqus_dloads)
        int jdl[16];
        double adl[16];
        int nd = 0;
        const int nloads = (int)m.dload_flag.size();
        for (int load = 0; load < nloads; ++load) {
            int es = m.dload_elset[load] - 1;
            for (int k = m.elementset_ptr[es]; k < m.elementset_ptr[es + 1]; ++k) {
                int e = m.elementset_elements[k] - 1;
                if (m.element_flag[e] < FLAG_UEL_BASE) break;       // :848, :876 "exit"
                if (e != lmn) continue;
                double bc_val = m.dload_values[m.dload_val_ptr[load]];
                int h = m.dload_history[load];
                if (h > 0) {
                    int o = m.history_ptr[h - 1], nh = m.history_ptr[h] - o;
                    double v1 = interpolate_history_table(&m.history_time[o], &m.history_value[o], nh, time + dtime);
                    bc_val = bc_val * v1;                            // :886-890
                }
                if (nd < 16) { jdl[nd] = m.dload_flag[load]; adl[nd] = bc_val; ++nd; }
            }
        }
        const int jtype = flag - FLAG_UEL_BASE;
        if (jtype == 1) {
            double energy[8], pnewdt;
            uel_linelastic_3d(xc, U, props, m.element_n_states[lmn], nd, jdl, adl, K, R, svars, energy, &pnewdt);
            return 0;
        } else if (jtype == 2) {
            int fail = 0;
            el_neohooke_3d(xc, U, props, K, R, svars, &fail);
            return fail;
        }
        std::fprintf(stderr, "oracle: unknown UEL type U%d\n", jtype);
        return 1;
    }
    if (flag == ID_LINELAST) {                                       // user_element.f90:62-69
        el_linelast_3dbasic(xc, du, ut, props, K, R);
        return 0;
    }
    if (flag == ID_NEOHOOKE) {
        double U[24];
        int fail = 0;
        for (int k = 0; k < 24; ++k) U[k] = du[k] + ut[k];
        el_neohooke_3d(xc, U, props, K, R, svars, &fail);
        return fail;
    }
    std::fprintf(stderr, "oracle: invalid element type %d (user_element.f90:83-86)\n", flag);
    return 1;
}

}  // namespace oracle
