// ORACLE — test infrastructure only (see or_model.h).  State print: projection of Gauss-point fields to the nodes and
// the Tecplot FEPOINT file (src/printing.f90:249-1093, :1096-1212, :1386-1668; element parts
// user_codes/src/el_linelast_3Dbasic.f90:266-413, user_codes/src/abaqus_tecplot_interface.f90:8-185,
// src/continuum_elements.f90:1206-1351).  3-D hex8 zones only, one zone holding every element.
// PARITY PINNING: no state-print output ships with the reference (Output_Files/ holds one log only) and the reference
// cannot be run here: "parity unpinned" for the file contents; the projection matrix itself is pinned through the
// conditioning line of the golden log (projection_mass_lu, or_direct.cpp).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "or_model.h"

namespace oracle {

static const double SGN[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};

// src/parse.f90:155-176: first `n` characters, either case
static bool strcmp_ref(const std::string& a, const char* b, int n) {
    for (int i = 0; i < n; ++i) {
        const int ia = i < (int)a.size() ? (unsigned char)a[i] : ' ', ib = (unsigned char)b[i];
        if (ia != ib && ia + 32 != ib && ia != ib + 32) return false;
    }
    return true;
}

static void shape(const double xi[3], double N[8], double dN[8][3]) {      // Element_Utilities.f90:908-940
    for (int a = 0; a < 8; ++a) {
        const double f1 = 1.0 + SGN[a][0] * xi[0], f2 = 1.0 + SGN[a][1] * xi[1], f3 = 1.0 + SGN[a][2] * xi[2];
        N[a] = f1 * f2 * f3 / 8.0;
        dN[a][0] = SGN[a][0] * (f2 * f3 / 8.0); dN[a][1] = SGN[a][1] * (f1 * f3 / 8.0); dN[a][2] = SGN[a][2] * (f1 * f2 / 8.0);
    }
}
static void jacobian(const double x[8][3], const double dN[8][3], double J[3][3]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { double t = 0; for (int a = 0; a < 8; ++a) t += x[a][r] * dN[a][c]; J[r][c] = t; }
}
static double det3(const double J[3][3]) {
    return J[0][0] * J[1][1] * J[2][2] - J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] +
           J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] - J[0][2] * J[1][1] * J[2][0];
}
static void inverse3(const double A[3][3], double d, double Ai[3][3]) {    // Element_Utilities.f90:99-123
    double C[3][3];
    C[0][0] = +(A[1][1] * A[2][2] - A[1][2] * A[2][1]); C[0][1] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]);
    C[0][2] = +(A[1][0] * A[2][1] - A[1][1] * A[2][0]); C[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]);
    C[1][1] = +(A[0][0] * A[2][2] - A[0][2] * A[2][0]); C[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]);
    C[2][0] = +(A[0][1] * A[1][2] - A[0][2] * A[1][1]); C[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]);
    C[2][2] = +(A[0][0] * A[1][1] - A[0][1] * A[1][0]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ai[i][j] = C[j][i] / d;
}
static double mises(const double s[6]) {
    const double p = (s[0] + s[1] + s[2]) / 3.0;
    const double d0 = s[0] - p, d1 = s[1] - p, d2 = s[2] - p;
    return std::sqrt((d0 * d0 + d1 * d1 + d2 * d2) + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5])) * std::sqrt(1.5);
}

// value of field `name` at an integration point; kind 0: 1001 / UEL (stress + Mises), 1: C3D8 (stress, strain, U<n>)
static bool point_value(const std::string& name, int kind, const double stress[6], const double strain[6], const double* sv,
                        int nsp, double* v) {
    static const char* S[6] = {"S11", "S22", "S33", "S12", "S13", "S23"};
    static const char* Ee[6] = {"E11", "E22", "E33", "E12", "E13", "E23"};
    for (int i = 0; i < 6; ++i)
        if (strcmp_ref(name, S[i], 3)) { *v = stress[i]; return true; }
    if (kind == 0) {
        if (strcmp_ref(name, "SMISES", 6)) { *v = mises(stress); return true; }
        return false;
    }
    for (int i = 0; i < 6; ++i)
        if (strcmp_ref(name, Ee[i], 3)) { *v = strain[i]; return true; }
    if (strcmp_ref(name, "U", 1)) {                                   // continuum_elements.f90:1331-1345
        size_t i = 1;
        while (i < name.size() && name[i] >= '0' && name[i] <= '9') ++i;
        if (i > 1) {
            const int nvar = std::atoi(name.substr(1, i - 1).c_str());
            if (nvar < nsp - 15) { *v = sv[14 + nvar]; return true; }
        }
    }
    return false;
}

int assemble_field_projection(const Model& m, const double* dof_total, const double* dof_increment, const double* svars,
                              const std::vector<std::string>& names, std::vector<double>& field) {
    const int nf = (int)names.size(), N = m.n_nodes;
    field.assign((size_t)nf * N, 0.0);
    const double g = (double)0.5773502692f;
    std::vector<double> nodal((size_t)nf * 8);
    size_t sv_off = 0;
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        const int flag = m.element_flag[lmn], ns = m.element_n_states[lmn];
        const double* sv = svars ? svars + sv_off : nullptr;
        sv_off += ns;
        double x[8][3], ut[24], du[24];
        int nodes[8];
        for (int a = 0; a < 8; ++a) {
            nodes[a] = m.connectivity[8 * (size_t)lmn + a] - 1;
            for (int i = 0; i < 3; ++i) {
                x[a][i] = m.coords[3 * (size_t)nodes[a] + i];
                ut[3 * a + i] = dof_total[3 * (size_t)nodes[a] + i];
                du[3 * a + i] = dof_increment[3 * (size_t)nodes[a] + i];
            }
        }
        std::fill(nodal.begin(), nodal.end(), 0.0);
        int kind;
        if (flag == ID_LINELAST) kind = 0;
        else if (flag == FLAG_UEL_BASE + 1) { kind = 0; if (!sv || ns < 48) return 2; }   // abaqus_tecplot_interface.f90:148-155
        else if (flag == FLAG_C3D8) { kind = 1; if (!sv || ns < 120) return 2; }
        else return 1;                                                                    // user_element.f90:258-266 stops
        const int nsp = ns / 8;
        double d11 = 0, d12 = 0, d44 = 0;
        if (flag == ID_LINELAST) {
            const double* pr = m.element_properties.data() + m.element_prop_index[lmn];
            const double E = pr[0], xnu = pr[1];
            d44 = 0.5 * E / (1 + xnu);
            d11 = (1.0 - xnu) * E / ((1 + xnu) * (1 - 2.0 * xnu));
            d12 = xnu * E / ((1 + xnu) * (1 - 2.0 * xnu));
        }
        for (int kint = 0; kint < 8; ++kint) {
            const double xi[3] = {(kint & 1) ? g : -g, (kint & 2) ? g : -g, (kint & 4) ? g : -g};
            double Nn[8], dN[8][3], J[3][3];
            shape(xi, Nn, dN);
            jacobian(x, dN, J);
            const double det = det3(J);
            double stress[6] = {0, 0, 0, 0, 0, 0}, strain[6] = {0, 0, 0, 0, 0, 0};
            const double* svp = nullptr;
            if (flag == ID_LINELAST) {                                  // el_linelast_3Dbasic.f90:369-391
                double Ji[3][3], dNdx[8][3];
                inverse3(J, det, Ji);
                for (int a = 0; a < 8; ++a)
                    for (int j = 0; j < 3; ++j) { double t = 0; for (int k = 0; k < 3; ++k) t += dN[a][k] * Ji[k][j]; dNdx[a][j] = t; }
                double e[6] = {0, 0, 0, 0, 0, 0}, de[6] = {0, 0, 0, 0, 0, 0};
                for (int pass = 0; pass < 2; ++pass) {
                    const double* u = pass ? du : ut;
                    double* q = pass ? de : e;
                    for (int a = 0; a < 8; ++a) {                       // strain = matmul(B, u): columns in DOF order
                        q[0] += dNdx[a][0] * u[3 * a];
                        q[1] += dNdx[a][1] * u[3 * a + 1];
                        q[2] += dNdx[a][2] * u[3 * a + 2];
                        q[3] += dNdx[a][1] * u[3 * a]; q[3] += dNdx[a][0] * u[3 * a + 1];
                        q[4] += dNdx[a][2] * u[3 * a]; q[4] += dNdx[a][0] * u[3 * a + 2];
                        q[5] += dNdx[a][2] * u[3 * a + 1]; q[5] += dNdx[a][1] * u[3 * a + 2];
                    }
                }
                double t[6];
                for (int i = 0; i < 6; ++i) t[i] = e[i] + de[i];
                stress[0] = d11 * t[0] + d12 * t[1] + d12 * t[2];
                stress[1] = d12 * t[0] + d11 * t[1] + d12 * t[2];
                stress[2] = d12 * t[0] + d12 * t[1] + d11 * t[2];
                stress[3] = d44 * t[3]; stress[4] = d44 * t[4]; stress[5] = d44 * t[5];
            } else if (kind == 0) {                                     // UEL: SVARS(6*kint-5:6*kint)
                for (int i = 0; i < 6; ++i) stress[i] = sv[6 * kint + i];
            } else {                                                    // C3D8: stress, strain per point
                svp = sv + (size_t)nsp * kint;
                for (int i = 0; i < 6; ++i) { stress[i] = svp[i]; strain[i] = svp[6 + i]; }
            }
            for (int k = 0; k < nf; ++k) {
                double v;
                if (!point_value(names[k], kind, stress, strain, svp, nsp, &v)) continue;
                for (int a = 0; a < 8; ++a) nodal[k + (size_t)nf * a] = nodal[k + (size_t)nf * a] + v * Nn[a] * det * 1.0;
            }
        }
        for (int a = 0; a < 8; ++a)
            for (int k = 0; k < nf; ++k) field[k + (size_t)nf * nodes[a]] = field[k + (size_t)nf * nodes[a]] + nodal[k + (size_t)nf * a];
    }
    return 0;
}

void lumped_projection_mass(const Model& m, std::vector<double>& lump) {
    lump.assign(m.n_nodes, 0.0);
    const double x1[3] = {(double)-0.7745966692f, 0.0, (double)0.7745966692f};
    const double w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        double x[8][3];
        int nodes[8];
        for (int a = 0; a < 8; ++a) {
            nodes[a] = m.connectivity[8 * (size_t)lmn + a] - 1;
            for (int i = 0; i < 3; ++i) x[a][i] = m.coords[3 * (size_t)nodes[a] + i];
        }
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) {
                    const double xi[3] = {x1[i], x1[j], x1[k]}, w = w1[i] * w1[j] * w1[k];
                    double Nn[8], dN[8][3], J[3][3];
                    shape(xi, Nn, dN);
                    jacobian(x, dN, J);
                    const double det = det3(J);
                    double sumN = 0.0;
                    for (int a = 0; a < 8; ++a) sumN += Nn[a];
                    for (int a = 0; a < 8; ++a) lump[nodes[a]] = lump[nodes[a]] + Nn[a] * sumN * det * w;
                }
    }
}

int project_fields(const Model& m, const double* dof_total, const double* dof_increment, const double* svars,
                   const std::vector<std::string>& names, bool lumped, std::vector<double>& field) {
    const int nf = (int)names.size(), N = m.n_nodes;
    if (int rc = assemble_field_projection(m, dof_total, dof_increment, svars, names, field)) return rc;
    if (nf == 0) return 0;
    if (lumped) {                                                       // printing.f90:464-471
        std::vector<double> lump;
        lumped_projection_mass(m, lump);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < nf; ++k) field[k + (size_t)nf * n] = field[k + (size_t)nf * n] / lump[n];
        return 0;
    }
    std::vector<int> order;                                             // printing.f90:475-488 (generate_node_numbers = GPS)
    gps_node_order(m, order);
    std::vector<int> num(N);
    for (int i = 0; i < N; ++i) num[order[i]] = i;
    Skyline s;
    projection_mass_factor(m, order, s, nullptr);
    s.rhs.assign(N, 0.0);
    for (int k = 0; k < nf; ++k) {
        for (int n = 0; n < N; ++n) s.rhs[num[n]] = field[k + (size_t)nf * n];
        skyline_backsubstitution(s);
        for (int n = 0; n < N; ++n) field[k + (size_t)nf * n] = s.rhs[num[n]];
    }
    return 0;
}

// Fortran E18.10 edit descriptor: 0.dddddddddd scaled mantissa, two-digit exponent (three digits drop the letter)
static void fmt_e18_10(double v, std::string& out) {
    char buf[64];
    if (!std::isfinite(v)) { std::snprintf(buf, sizeof buf, "%18s", std::isnan(v) ? "NaN" : (v > 0 ? "Infinity" : "-Infinity")); out += buf; return; }
    std::snprintf(buf, sizeof buf, "%.9E", std::fabs(v));              // d.dddddddddE+xx : 10 significant digits
    char digits[16];
    int nd = 0;
    for (const char* p = buf; *p && *p != 'E'; ++p)
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
    digits[nd] = 0;
    int ex = std::atoi(std::strchr(buf, 'E') + 1) + 1;
    if (v == 0.0) ex = 0;
    char e[16];
    if (std::abs(ex) < 100) std::snprintf(e, sizeof e, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
    else std::snprintf(e, sizeof e, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
    std::string body = std::string(std::signbit(v) ? "-" : "") + "0." + digits + e;
    if (body.size() < 18) out.append(18 - body.size(), ' ');
    out += body;
}
static void fmt_d12_4(double v, std::string& out) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.3E", std::fabs(v));
    char digits[8];
    int nd = 0;
    for (const char* p = buf; *p && *p != 'E'; ++p)
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
    digits[nd] = 0;
    int ex = std::atoi(std::strchr(buf, 'E') + 1) + 1;
    if (v == 0.0) ex = 0;
    char e[16];
    std::snprintf(e, sizeof e, "D%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
    std::string body = std::string(std::signbit(v) ? "-" : "") + "0." + digits + e;
    if (body.size() < 12) out.append(12 - body.size(), ' ');
    out += body;
}
static void fmt_i5(long long v, std::string& out) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%5lld", v);
    out += (std::strlen(buf) > 5) ? "*****" : buf;                     // Fortran I5 overflow
}

int print_state(const Model& m, const PrintOptions& o, double time_end, const double* dof_total,
                const double* dof_increment, const double* svars, std::string& out) {
    const int N = m.n_nodes, nf = (int)o.field_names.size();
    std::vector<double> field;
    if (nf > 0)
        if (int rc = project_fields(m, dof_total, dof_increment, svars, o.field_names, o.lumped, field)) return rc;
    // printing.f90:358-388
    std::string head = "VARIABLES = X,Y,Z";
    if (o.print_dof) {
        for (int i = 1; i <= 3; ++i) head += ",DU" + std::to_string(i);
        for (int i = 1; i <= 3; ++i) head += ",U" + std::to_string(i);
    }
    for (const std::string& nm : o.field_names) head += "," + nm;
    out += " " + head + "\n";
    // printable node numbers: first appearance in element order (:800-825)
    std::vector<int> num(N, 0);
    int np = 0;
    for (int lmn = 0; lmn < m.n_elements; ++lmn)
        for (int k = 0; k < 8; ++k) {
            const int n = m.connectivity[8 * (size_t)lmn + k] - 1;
            if (num[n] == 0) num[n] = ++np;
        }
    out += " ZONE, T=\"";
    fmt_d12_4(time_end, out);
    out += "\" F=FEPOINT, I=";
    fmt_i5(np, out);
    out += " J=";
    fmt_i5(m.n_elements, out);
    out += " ET=BRICK\n";
    std::vector<char> done(N, 0);
    for (int lmn = 0; lmn < m.n_elements; ++lmn)
        for (int k = 0; k < 8; ++k) {
            const int n = m.connectivity[8 * (size_t)lmn + k] - 1;
            if (done[n]) continue;
            done[n] = 1;
            double xx[3];
            for (int i = 0; i < 3; ++i) {
                xx[i] = m.coords[3 * (size_t)n + i];
                if (o.displaced_mesh) xx[i] = xx[i] + o.displacement_scale_factor * (dof_total[3 * (size_t)n + i] + dof_increment[3 * (size_t)n + i]);
            }
            for (int i = 0; i < 3; ++i) { out += ' '; fmt_e18_10(xx[i], out); }
            if (o.print_dof) {
                for (int i = 0; i < 3; ++i) { out += ' '; fmt_e18_10(dof_increment[3 * (size_t)n + i], out); }
                for (int i = 0; i < 3; ++i) { out += ' '; fmt_e18_10(dof_total[3 * (size_t)n + i], out); }
            }
            for (int i = 0; i < nf; ++i) { out += ' '; fmt_e18_10(field[i + (size_t)nf * n], out); }
            out += '\n';
        }
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {                      // (8(1x,i5)) :742-744
        for (int k = 0; k < 8; ++k) { out += ' '; fmt_i5(num[m.connectivity[8 * (size_t)lmn + k] - 1], out); }
        out += '\n';
    }
    return 0;
}

}  // namespace oracle
