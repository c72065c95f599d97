// ORACLE — test infrastructure only (see or_model.h).  SURVEY 8(f) rank-1 row: the built-in C3D8 B-bar finite-strain
// continuum element with the linear-elastic (hypoelastic) ABAQUS UMAT.  It is NOT on the GPU hot path; it is restated
// here only because the single golden output of the reference, Output_Files/Abaqus_umat_holeplate_3d.out, exercises it:
// reproducing that log pins the oracle's driver / assembly / BC / skyline-LDU / Newton logic with reference output.
//   element: src/continuum_elements.f90:368-761 (continuum_element_static_3D), hex8 branch
//   material: user_codes/src/abaqus_umat_elastic.for:253-308
#include <algorithm>
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

// ---------------------------------------------------------------------------------------------------------------
// Gibbs-Poole-Stockmeyer node renumbering, src/bandwidth_minimizer.f90 (generate_node_numbers :2-128 and its helpers),
// restated with the reference's tie-breaking so that the skyline profile of the golden log is reproduced exactly:
// adjacency lists in insertion order (compute_adjacency :250-268; the linked 10-entry bins of
// modules/Linkedlist_Handling.f90 only preserve insertion order, so plain vectors are equivalent), the heapsort index
// routine isort :1084-1138 (its treatment of equal keys matters), level structures in insertion order (levgen
// :1001-1082), pseudo-diameter search diam :335-484, level-width minimisation widmin :486-736, numbering numgen
// :738-931 with lowfnd :933-999.  1-based node numbers inside, like the reference.  No constraints (n_constraints = 0).
namespace {

// isort :1084-1138 — index(1..n) lists the entries of `in` in increasing magnitude (heapsort on the index table)
void isort(int n, const std::vector<int>& in /*1-based*/, std::vector<int>& index /*1-based*/) {
    if ((int)index.size() < n + 2) index.resize(n + 2);
    for (int j = 1; j <= n; ++j) index[j] = j;
    if (n <= 1) return;
    int l = n / 2 + 1, ir = n, indxt, k;
    for (;;) {
        if (l > 1) {
            l = l - 1;
            indxt = index[l];
            k = in[indxt];
        } else {
            indxt = index[ir];
            k = in[indxt];
            index[ir] = index[1];
            ir = ir - 1;
            if (ir == 1) { index[1] = indxt; return; }
        }
        int i = l, j = l + l;
        while (j <= ir) {
            if (j < ir && in[index[j]] < in[index[j + 1]]) j = j + 1;
            if (k < in[index[j]]) { index[i] = index[j]; i = j; j = j + j; }
            else j = ir + 1;
        }
        index[i] = indxt;
    }
}

struct Levels {
    std::vector<int> node_level;                 // 1-based node -> level (0 = none)
    std::vector<std::vector<int>> list;          // list[l], l = 1..n_levels, nodes in insertion order
    int n_levels = 0, width = 0;
    int w(int l) const { return l < (int)list.size() ? (int)list[l].size() : 0; }
};

struct Gps {
    int nops = 0;
    std::vector<int> ideg;                       // node_degree (1-based)
    std::vector<std::vector<int>> adj;           // adjacency in insertion order (1-based nodes)

    // levgen :1001-1082 with the degree array passed in (widmin calls it with the reduced degrees)
    void levgen(const std::vector<int>& deg, int start, Levels& L) const {
        L.node_level.assign(nops + 1, 0);
        L.list.assign(2, {});
        L.node_level[start] = 1;
        L.list[1].push_back(start);
        L.n_levels = 1;
        for (;;) {
            const int nl1 = L.n_levels;
            L.n_levels++;
            if (L.n_levels > nops) break;
            L.list.emplace_back();
            const std::vector<int> prev = L.list[nl1];
            for (int node : prev)
                for (int j = 0; j < deg[node]; ++j) {
                    const int nbr = adj[node][j];
                    if (L.node_level[nbr] == 0 && deg[nbr] > 0) {
                        L.list[L.n_levels].push_back(nbr);
                        L.node_level[nbr] = L.n_levels;
                    }
                }
            if (L.list[L.n_levels].empty()) break;
        }
        L.n_levels--;
        L.width = 0;
        for (int l = 1; l <= L.n_levels; ++l) L.width = std::max(L.width, (int)L.list[l].size());
    }
};

}  // namespace

void gps_node_order(const Model& m, std::vector<int>& node_order) {
    const int N = m.n_nodes;
    Gps g;
    g.nops = N;
    g.ideg.assign(N + 1, 0);
    g.adj.assign(N + 1, {});
    // compute_adjacency :131-333 (elements only): adjacency in order of first appearance
    for (int lmn = 0; lmn < m.n_elements; ++lmn)
        for (int i = 0; i < 8; ++i) {
            const int node1 = m.connectivity[8 * (size_t)lmn + i];
            for (int j = 0; j < 8; ++j) {
                const int node2 = m.connectivity[8 * (size_t)lmn + j];
                if (node2 == node1) continue;
                auto& a = g.adj[node1];
                if (std::find(a.begin(), a.end(), node2) == a.end()) a.push_back(node2);
            }
        }
    for (int n = 1; n <= N; ++n) g.ideg[n] = (int)g.adj[n].size();

    // ---- diam :335-484
    int mindeg = N, minnod = 1;
    for (int i = 1; i <= N; ++i)
        if (g.ideg[i] < mindeg && g.ideg[i] > 0) { mindeg = g.ideg[i]; minnod = i; }
    int nv = minnod, nu = minnod;
    Levels V, U, S;
    g.levgen(g.ideg, nv, V);
    std::vector<int> id, sidx;
    bool swapped = true;
    while (swapped) {
        const std::vector<int> last = V.list[V.n_levels];
        const int n = (int)last.size();
        id.assign(n + 2, 0);
        for (int i = 1; i <= n; ++i) id[i] = g.ideg[last[i - 1]];
        isort(n, id, sidx);
        int minrem = N;
        swapped = false;
        for (int i = 1; i <= n; ++i) {
            const int ns = last[sidx[i] - 1];
            g.levgen(g.ideg, ns, S);
            swapped = false;
            if (S.n_levels > V.n_levels) { nv = ns; V = S; swapped = true; break; }
            if (S.width <= minrem) { minrem = S.width; nu = ns; U = S; }
        }
    }
    const int iwvmax = V.width, iwumax = U.width;

    // ---- widmin :486-736
    Levels F;                                                   // final level structure
    F.node_level.assign(N + 1, 0);
    F.list.assign(N + 2, {});
    int ipair = 0;
    {
        int icount = 0;
        for (int i = 1; i <= N; ++i) if (g.ideg[i] > 0) icount++;
        std::vector<int> ideg2 = g.ideg;
        for (int i = 1; i <= N; ++i)
            if (g.ideg[i] > 0 && V.node_level[i] == U.n_levels + 1 - U.node_level[i]) {
                F.node_level[i] = V.node_level[i];
                F.list[V.node_level[i]].push_back(i);
                icount--;
                ideg2[i] = 0;
            }
        if (icount != 0) {
            std::vector<std::vector<int>> comp;
            for (int i = 1; i <= N; ++i)
                if (ideg2[i] != 0) {
                    g.levgen(ideg2, i, S);
                    comp.emplace_back();
                    for (int l = 1; l <= S.n_levels; ++l)
                        for (int node : S.list[l]) { comp.back().push_back(node); ideg2[node] = 0; }
                }
            const int ncon = (int)comp.size();
            std::vector<int> nmem(ncon + 2, 0), index;
            for (int c = 1; c <= ncon; ++c) nmem[c] = (int)comp[c - 1].size();
            isort(ncon, nmem, index);
            std::vector<int> levwh(N + 2), levwl(N + 2), lw(N + 2);
            for (int i = 1; i <= ncon; ++i) {
                const std::vector<int>& cn = comp[index[ncon + 1 - i] - 1];
                for (int l = 1; l <= N; ++l) lw[l] = (int)F.list[l].size();
                levwh = lw; levwl = lw;
                for (int node : cn) { levwh[V.node_level[node]]++; levwl[U.n_levels + 1 - U.node_level[node]]++; }
                int ihmax = 0, ilmax = 0;
                for (int j = 1; j <= U.n_levels; ++j) {
                    if (levwh[j] > ihmax && levwh[j] > lw[j]) ihmax = levwh[j];
                    if (levwl[j] > ilmax && levwl[j] > lw[j]) ilmax = levwl[j];
                }
                const bool use_v = ilmax > ihmax ? true : (ilmax < ihmax ? false : iwvmax <= iwumax);
                if (i == 1) ipair = use_v ? 1 : 2;
                for (int node : cn) {
                    const int ilev = use_v ? V.node_level[node] : U.n_levels + 1 - U.node_level[node];
                    F.node_level[node] = ilev;
                    F.list[ilev].push_back(node);
                }
            }
        }
    }
    const int n_levels = V.n_levels;

    // ---- numgen :738-931
    std::vector<int> numnod(N + 1, 0), list(N + 2), ldeg(N + 2), sort_index;
    int irev;
    if (g.ideg[nv] <= g.ideg[nu]) { numnod[nv] = 1; irev = 0; } else { numnod[nu] = 1; irev = 1; }
    int number = 1;
    // lowfnd :933-999: lowest-numbered node of lev1 with unnumbered neighbours in lev2; their list in adjacency order
    auto lowfnd = [&](const std::vector<int>& lev1, const std::vector<int>& lev2) -> int {
        int low = N + 1, lcount = 0;
        for (int node1 : lev1) {
            if (numnod[node1] < low && numnod[node1] > 0) {
                std::vector<int> tmp;
                for (int nbr : g.adj[node1])
                    if (numnod[nbr] == 0 && std::find(lev2.begin(), lev2.end(), nbr) != lev2.end()) tmp.push_back(nbr);
                if (!tmp.empty()) {
                    low = numnod[node1];
                    lcount = (int)tmp.size();
                    for (int j = 1; j <= lcount; ++j) { list[j] = tmp[j - 1]; ldeg[j] = g.ideg[list[j]]; }
                }
            }
        }
        return lcount;
    };
    auto number_list = [&](int lcount) {
        isort(lcount, ldeg, sort_index);
        for (int i = 1; i <= lcount; ++i) numnod[list[sort_index[i]]] = ++number;
    };
    {   // first (last) level
        const int l = irev == 1 ? n_levels : 1;
        const std::vector<int>& lev = F.list[l];
        for (;;) {
            const int lcount = lowfnd(lev, lev);
            if (lcount > 0) { number_list(lcount); continue; }
            int nodrem = 0, md = N;
            for (int node : lev)
                if (numnod[node] == 0 && g.ideg[node] < md && g.ideg[node] > 0) { md = g.ideg[node]; nodrem = node; }
            if (nodrem > 0) { numnod[nodrem] = ++number; continue; }
            break;
        }
    }
    {
        const int lstart = irev == 0 ? 2 : n_levels - 1, lend = irev == 0 ? n_levels : 1, linc = irev == 0 ? 1 : -1;
        for (int l = lstart; irev == 0 ? l <= lend : l >= lend; l += linc) {
            const std::vector<int>& lev1 = F.list[l - linc];
            const std::vector<int>& lev2 = F.list[l];
            for (;;) {
                const int lcount = lowfnd(lev1, lev2);
                if (lcount > 0) { number_list(lcount); continue; }
                int nodrem = 0, md = N;
                for (int node : lev2)
                    if (numnod[node] == 0 && g.ideg[node] <= md && g.ideg[node] > 0) { md = g.ideg[node]; nodrem = node; }
                if (nodrem > 0) { numnod[nodrem] = ++number; continue; }
                break;
            }
        }
    }
    int nodrem = 0;
    for (int i = 1; i <= N; ++i) if (g.ideg[i] > 0) nodrem++;
    if ((irev == 1 && ipair == 2) || (irev == 0 && ipair == 1))
        for (int i = 1; i <= N; ++i) if (g.ideg[i] > 0) numnod[i] = nodrem + 1 - numnod[i];
    // node_order_index = isort(node_numbers) (:102): position -> node
    std::vector<int> order_index;
    isort(N, numnod, order_index);
    node_order.resize(N);
    for (int i = 1; i <= N; ++i) node_order[i - 1] = order_index[i] - 1;
}

}  // namespace oracle
