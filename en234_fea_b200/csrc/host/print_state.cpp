// print_state of the product (src/printing.f90:249-1093): the field projection runs on the device
// (en234gpu_project_fields), this file names the fields and writes the Tecplot text.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "print_state.h"

namespace en234 {
namespace {

bool prefix(const std::string& name, const char* key) {             // strcmp of src/parse.f90:155-176
    const size_t len = std::strlen(key);
    for (size_t i = 0; i < len; ++i) {
        const int a = i < name.size() ? (unsigned char)name[i] : ' ', b = (unsigned char)key[i];
        if (a != b && a + 32 != b && a != b + 32) return false;
    }
    return true;
}

// Fortran Ew.d / Dw.d output: 0.ddd mantissa, sign, two-digit exponent (three digits drop the letter), right-justified
void fortran_exp(std::string& out, double v, int width, int digits, char letter) {
    std::string body;
    if (std::isnan(v)) body = "NaN";
    else if (std::isinf(v)) body = v > 0 ? "Infinity" : "-Infinity";
    else {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.*e", digits - 1, std::fabs(v));
        const char* epos = std::strchr(buf, 'e');
        int ex = v == 0.0 ? 0 : std::atoi(epos + 1) + 1;
        body = std::signbit(v) ? "-0." : "0.";
        for (const char* p = buf; p < epos; ++p)
            if (*p != '.') body += *p;
        char e[16];
        if (std::abs(ex) < 100) std::snprintf(e, sizeof e, "%c%c%02d", letter, ex < 0 ? '-' : '+', std::abs(ex));
        else std::snprintf(e, sizeof e, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
        body += e;
    }
    if ((int)body.size() < width) out.append(width - body.size(), ' ');
    out += body;
}
// I5; the reference prints ***** above 99999 — the count is written in full here so that large meshes stay readable
void int5(std::string& out, long long v) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%5lld", v);
    out += buf;
}

}  // namespace

std::string normalise_path(const std::string& s) {
    std::string t = s;
    for (char& c : t)
        if (c == '\\') c = '/';
    return t;
}

std::vector<int> field_codes(const Model& m, const std::vector<std::string>& names) {
    const bool c3d8 = !m.element_flag.empty() && m.element_flag[0] == EN234_FLAG_C3D8;
    const int nsp = (c3d8 && !m.element_n_states.empty()) ? m.element_n_states[0] / 8 : 0;
    static const char* S[6] = {"S11", "S22", "S33", "S12", "S13", "S23"};
    static const char* E[6] = {"E11", "E22", "E33", "E12", "E13", "E23"};
    std::vector<int> codes;
    for (const std::string& nm : names) {
        int code = EN234_FIELD_NONE;
        for (int i = 0; i < 6 && code == EN234_FIELD_NONE; ++i)
            if (prefix(nm, S[i])) code = EN234_FIELD_S11 + i;
        if (code == EN234_FIELD_NONE && !c3d8 && prefix(nm, "SMISES")) code = EN234_FIELD_SMISES;
        if (code == EN234_FIELD_NONE && c3d8) {
            for (int i = 0; i < 6 && code == EN234_FIELD_NONE; ++i)
                if (prefix(nm, E[i])) code = EN234_FIELD_E11 + i;
            if (code == EN234_FIELD_NONE && prefix(nm, "U")) {       // U<n>: material state variable n of the point
                size_t i = 1;
                while (i < nm.size() && nm[i] >= '0' && nm[i] <= '9') ++i;
                if (i > 1) {
                    const int nvar = std::atoi(nm.substr(1, i - 1).c_str());
                    if (nvar < nsp - 15) code = EN234_FIELD_STATE0 + 14 + nvar;
                }
            }
        }
        codes.push_back(code);
    }
    return codes;
}

int project_fields(const Model& m, en234gpu_handle h, const std::vector<std::string>& names, bool lumped,
                   const double* dof_total, const double* dof_increment, std::vector<double>& fields, std::string& err) {
    const int nf = (int)names.size();
    fields.assign((size_t)nf * m.n_nodes, 0.0);
    if (nf == 0) return 0;
    if (nf > EN234_MAX_FIELDS) { err = "at most " + std::to_string(EN234_MAX_FIELDS) + " FIELD VARIABLES"; return 1; }
    for (int f : m.element_flag)
        if (f != 1001 && f != 100000 && f != EN234_FLAG_C3D8) {      // user_element.f90:258-266 stops here
            err = "no field variables are defined for element type " + std::to_string(f);
            return 1;
        }
    const std::vector<int> codes = field_codes(m, names);
    int its = 0;
    if (en234gpu_project_fields(h, nf, codes.data(), lumped ? 1 : 0, 0.0, 0, dof_total, dof_increment, fields.data(), &its)) {
        err = en234gpu_last_error();
        return 1;
    }
    return 0;
}

int print_state(const Model& m, en234gpu_handle h, double time_end, const double* dof_total, const double* dof_increment,
                FILE* f, std::string& err) {
    std::vector<double> fields;
    if (project_fields(m, h, m.field_variable_names, m.use_lumped_projection_matrix, dof_total, dof_increment, fields, err)) return 1;
    write_state_zone(m, time_end, dof_total, dof_increment, fields, f);
    return 0;
}

// Partitioned run: every rank projects the fields on its part (the projection's right-hand sides and matrix are halo-summed,
// its solve is the partitioned PCG loop), the rank-local nodal values and DOFs are summed into whole-mesh arrays over the
// communicator together with a copy count (copies of a shared node agree to rounding: the mean is written), and the rank
// that holds the open file writes the zone.  Collective: every rank must call it.
int print_state_partitioned(const Model& m, en234gpu_handle h, double time_end, const double* dof_total_local,
                            const double* dof_increment_local, const std::vector<int>& g2l, size_t n_local_dof, FILE* f,
                            std::string& err) {
    const size_t N = (size_t)m.n_nodes, nloc = n_local_dof / 3;
    const int nf = (int)m.field_variable_names.size();
    if (nf > EN234_MAX_FIELDS) { err = "at most " + std::to_string(EN234_MAX_FIELDS) + " FIELD VARIABLES"; return 1; }
    for (int fl : m.element_flag)
        if (fl != 1001 && fl != 100000 && fl != EN234_FLAG_C3D8) { err = "no field variables are defined for element type " + std::to_string(fl); return 1; }
    std::vector<double> local((size_t)nf * nloc, 0.0);
    if (nf > 0) {
        const std::vector<int> codes = field_codes(m, m.field_variable_names);
        int its = 0;
        if (en234gpu_project_fields(h, nf, codes.data(), m.use_lumped_projection_matrix ? 1 : 0, 0.0, 0, dof_total_local,
                                    dof_increment_local, local.data(), &its)) { err = en234gpu_last_error(); return 1; }
    }
    // rows: nf fields, 3 dof_increment, 3 dof_total, 1 copy count
    std::vector<double> G((size_t)(nf + 7) * N, 0.0);
    for (size_t n = 0; n < N; ++n) {
        const int l = g2l[3 * n];
        if (l < 0) continue;
        const size_t ln = (size_t)l / 3;
        for (int k = 0; k < nf; ++k) G[(size_t)k * N + n] = local[(size_t)k * nloc + ln];
        for (int i = 0; i < 3; ++i) {
            G[(size_t)(nf + i) * N + n] = dof_increment_local[3 * ln + i];
            G[(size_t)(nf + 3 + i) * N + n] = dof_total_local[3 * ln + i];
        }
        G[(size_t)(nf + 6) * N + n] = 1.0;
    }
    if (en234gpu_comm_allreduce_host(h, G.data(), (long long)G.size())) { err = en234gpu_last_error(); return 1; }
    if (!f) return 0;
    std::vector<double> fields((size_t)nf * N), du(3 * N), ut(3 * N);
    for (size_t n = 0; n < N; ++n) {
        const double cnt = G[(size_t)(nf + 6) * N + n] > 0.0 ? G[(size_t)(nf + 6) * N + n] : 1.0;
        for (int k = 0; k < nf; ++k) fields[(size_t)k * N + n] = G[(size_t)k * N + n] / cnt;
        for (int i = 0; i < 3; ++i) {
            du[3 * n + i] = G[(size_t)(nf + i) * N + n] / cnt;
            ut[3 * n + i] = G[(size_t)(nf + 3 + i) * N + n] / cnt;
        }
    }
    write_state_zone(m, time_end, ut.data(), du.data(), fields, f);
    return 0;
}

void write_state_zone(const Model& m, double time_end, const double* dof_total, const double* dof_increment,
                      const std::vector<double>& fields, FILE* f) {
    const int N = m.n_nodes, nf = (int)m.field_variable_names.size();
    std::string out = " VARIABLES = X,Y,Z";                           // :358-388 (list-directed write: leading blank)
    if (m.print_dof) {
        for (int i = 1; i <= 3; ++i) out += ",DU" + std::to_string(i);
        for (int i = 1; i <= 3; ++i) out += ",U" + std::to_string(i);
    }
    for (const std::string& nm : m.field_variable_names) out += "," + nm;
    out += "\n";
    // printable node numbers in order of first appearance in the element list (:800-825)
    std::vector<int> number(N, 0), order;
    order.reserve(N);
    for (int c : m.connectivity)
        if (number[c - 1] == 0) { order.push_back(c - 1); number[c - 1] = (int)order.size(); }
    out += " ZONE, T=\"";                                             // (A10,D12.4,A15,I5,A3,I5,A9) :830-833
    fortran_exp(out, time_end, 12, 4, 'D');
    out += "\" F=FEPOINT, I=";
    int5(out, (long long)order.size());
    out += " J=";
    int5(out, m.n_elements);
    out += " ET=BRICK\n";
    std::fputs(out.c_str(), f);
    for (int n : order) {                                             // (n(1X,E18.10)) :869-895
        out.clear();
        for (int i = 0; i < 3; ++i) {
            double x = m.coords[3 * (size_t)n + i];
            if (m.print_displacedmesh) x = x + m.displacementscalefactor * (dof_total[3 * (size_t)n + i] + dof_increment[3 * (size_t)n + i]);
            out += ' ';
            fortran_exp(out, x, 18, 10, 'E');
        }
        if (m.print_dof) {
            for (int i = 0; i < 3; ++i) { out += ' '; fortran_exp(out, dof_increment[3 * (size_t)n + i], 18, 10, 'E'); }
            for (int i = 0; i < 3; ++i) { out += ' '; fortran_exp(out, dof_total[3 * (size_t)n + i], 18, 10, 'E'); }
        }
        for (int k = 0; k < nf; ++k) { out += ' '; fortran_exp(out, fields[(size_t)k * N + n], 18, 10, 'E'); }
        out += '\n';
        std::fputs(out.c_str(), f);
    }
    for (int e = 0; e < m.n_elements; ++e) {                          // (8(1x,i5)) :742-744
        out.clear();
        for (int k = 0; k < 8; ++k) { out += ' '; int5(out, number[m.connectivity[8 * (size_t)e + k] - 1]); }
        out += '\n';
        std::fputs(out.c_str(), f);
    }
    std::fflush(f);
}

}  // namespace en234
