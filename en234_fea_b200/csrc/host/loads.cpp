// Host-side evaluation of the boundary-condition data that feeds the device path each time the reference
// would call its assembly / BC routines.  These are surface (O(n^2)) terms, kept on the host:
//   * element_ext: ABAQUS-UEL distributed loads, which the reference's UEL adds to its own element residual
//     (user_codes/src/abaqus_uel_linelastic_3d.for:207-238; magnitudes from generate_abaqus_dloads,
//     src/solver_direct.f90:798-912)
//   * ext: native-element tractions (src/solver_direct.f90:501-568 + src/traction_boundarycondition.f90:66-94)
//     and prescribed nodal forces (src/solver_direct.f90:646-682)
//   * prescribed DOFs -> (dof, value) list in application order (src/solver_direct.f90:685-734)
#include <cmath>

#include "loads.h"

namespace en234 {
namespace {

// local node numbers (1-based) of the faces of an 8-node brick (Element_Utilities.f90:1102-1109)
const int FACE[6][4] = {{1, 2, 3, 4}, {5, 8, 7, 6}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 8, 4}, {4, 8, 5, 1}};

// 2x2 Gauss rule on a quadrilateral face; abscissa is a proper double here
// (Element_Utilities.f90:486 / abaqus_uel_2D.for:190), point order (-,-),(+,-),(+,+),(-,+)
const double CN = 0.5773502691896260;
const int GX[4] = {-1, 1, 1, -1}, GY[4] = {-1, -1, 1, 1};

struct FacePoint { double N[4]; double norm[3]; };

// shape functions of the 4-node face (abaqus_uel_2D.for:311-334) and the (unnormalised) normal as the
// reference forms it — including its sign convention for the second component (…linelastic_3d.for:228-230)
void face_point(const double fc[4][3], int kint, FacePoint& fp) {
    const double xi = GX[kint] * CN, eta = GY[kint] * CN;
    const double g1 = 0.5 * (1.0 - xi), g2 = 0.5 * (1.0 + xi), h1 = 0.5 * (1.0 - eta), h2 = 0.5 * (1.0 + eta);
    fp.N[0] = g1 * h1; fp.N[1] = g2 * h1; fp.N[2] = g2 * h2; fp.N[3] = g1 * h2;
    const double d1[4] = {-0.5 * h1, 0.5 * h1, 0.5 * h2, -0.5 * h2};
    const double d2[4] = {g1 * -0.5, g2 * -0.5, g2 * 0.5, g1 * 0.5};
    double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int a = 0; a < 4; ++a) { t1[i] += fc[a][i] * d1[a]; t2[i] += fc[a][i] * d2[a]; }
    fp.norm[0] = t1[1] * t2[2] - t2[1] * t1[2];
    fp.norm[1] = t1[0] * t2[2] - t2[0] * t1[2];
    fp.norm[2] = t1[0] * t2[1] - t2[0] * t1[1];
}

void face_coords(const Model& m, int lmn, int face, int nodes[4], double fc[4][3]) {
    for (int j = 0; j < 4; ++j) {
        nodes[j] = m.connectivity[8 * (size_t)lmn + FACE[face - 1][j] - 1] - 1;
        for (int i = 0; i < 3; ++i) fc[j][i] = m.coords[3 * (size_t)nodes[j] + i];
    }
}

}  // namespace

// DLOAD contribution of one ABAQUS-UEL face to the 24-entry element residual (abaqus_uel_linelastic_3d.for:207-238),
// coords24 = xyz per element node
void uel_face_load(const double* coords24, int face, double mag, double* R24) {
    double fc[4][3];
    int nodes[4];
    for (int j = 0; j < 4; ++j) {
        nodes[j] = FACE[face - 1][j] - 1;
        for (int i = 0; i < 3; ++i) fc[j][i] = coords24[3 * nodes[j] + i];
    }
    for (int kint = 0; kint < 4; ++kint) {
        FacePoint fp;
        face_point(fc, kint, fp);
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 3; ++c) R24[3 * nodes[i] + c] -= fp.N[c] * mag * fp.norm[c];
    }
}

// user_prescribeddof (user_codes/src/user_prescribeddof.f90:1-108) as shipped: uniaxial stretch of a bar over the first
// n_extension_steps steps, then a rigid rotation in n_rotation_steps steps, driven by current_step_number.
// xref / utot: reference coordinates and dof_total of the node (3 each).
static void user_prescribeddof(const Model& m, int node_number, int idof, const std::vector<double>& par, const double* xref,
                               const double* utot, double& dofvalue, bool& ignoredof) {
    if (par.size() < 11) { ignoredof = true; return; }
    const int tensile_dir = (int)par[0], n_ext = (int)par[2], rot_axis = (int)par[4], n_rot = (int)par[5];
    const int n_base = (int)par[6], n_end = (int)par[7], base_fixed = (int)par[8], base_con = (int)par[9], base_con_dir = (int)par[10];
    const double extension = par[1], rotation_angle = par[3];
    const int step = m.current_step_number;
    ignoredof = true;
    if (step <= n_ext) {
        for (int i = 11; i < 11 + n_base && i < (int)par.size(); ++i)
            if (node_number == (int)par[i] && idof == tensile_dir) { dofvalue = 0.0; ignoredof = false; }
        for (int i = 11 + n_base; i < 11 + n_base + n_end && i < (int)par.size(); ++i)
            if (node_number == (int)par[i] && idof == tensile_dir) { dofvalue = extension * step / n_ext; ignoredof = false; }
        if (node_number == base_fixed) { dofvalue = 0.0; ignoredof = false; }
        if (node_number == base_con && idof == base_con_dir) { dofvalue = 0.0; ignoredof = false; }
    } else if (step <= n_ext + n_rot) {
        const double PI_D = 3.141592653589793238462643383279502884197;
        double cur[3], axis[3] = {0, 0, 0}, cr[3];
        for (int i = 0; i < 3; ++i) cur[i] = xref[i] + utot[i];
        const double angle = PI_D * rotation_angle / (n_rot * 180.0);
        axis[rot_axis - 1] = 1.0;
        cr[0] = axis[1] * cur[2] - axis[2] * cur[1];                  // cross_product, Element_Utilities.f90:38-52
        cr[1] = cur[0] * axis[2] - axis[0] * cur[2];
        cr[2] = axis[0] * cur[1] - axis[1] * cur[0];
        const double ad = axis[0] * cur[0] + axis[1] * cur[1] + axis[2] * cur[2];
        const int i = idof - 1;
        dofvalue = std::cos(angle) * cur[i] + (1.0 - std::cos(angle)) * ad * axis[i] + std::sin(angle) * cr[i] - xref[i];
        ignoredof = false;
    } else { dofvalue = 0.0; ignoredof = false; }
}

void evaluate_loads(const Model& m, double time, double dtime, const double* dof_total, const double* dof_increment,
                    StepLoads& out, const std::vector<int>* g2l, size_t n_local) {
    // g2l (partitioned runs): global DOF -> rank-local DOF or -1; all outputs and the DOF arrays are then rank-local
    const size_t ndof = g2l ? n_local : (size_t)m.ndof();
    auto loc = [&](size_t gd) -> long { return g2l ? (long)(*g2l)[gd] : (long)gd; };
    out.element_ext.assign(ndof, 0.0);
    out.ext.assign(ndof, 0.0);
    out.fixed_dof.clear();
    out.fixed_value.clear();
    out.fixed_correction.clear();
    out.any_element_ext = out.any_ext = false;

    for (const auto& d : m.distributed_loads) {
        const NamedSet& es = m.elementsets[d.element_set - 1];
        double hist = 1.0;
        if (d.history_number > 0) hist = interpolate_history_table(m.histories[d.history_number - 1], time + dtime);
        for (int e1 : es.items) {
            const int lmn = e1 - 1;
            const bool uel = m.element_flag[lmn] > 99999;
            if (d.uel_format) {
                if (!uel) break;                                      // solver_direct.f90:848,:876 `exit`
                // pressure magnitude (:886-890): value * history(TIME+DTIME)
                const double mag = d.history_number > 0 ? d.values[0] * hist : d.values[0];
                int nodes[4];
                double fc[4][3];
                face_coords(m, lmn, d.flag < 0 ? -d.flag : d.flag, nodes, fc);
                for (int kint = 0; kint < 4; ++kint) {
                    FacePoint fp;
                    face_point(fc, kint, fp);
                    // abaqus_uel_linelastic_3d.for:232-236 as the compiler scalarises it: every face
                    // node receives (N2(1) n1, N2(2) n2, N2(3) n3)
                    for (int i = 0; i < 4; ++i)
                        for (int c = 0; c < 3; ++c) {
                            const long ld = loc(3 * (size_t)nodes[i] + c);
                            if (ld >= 0) out.element_ext[ld] -= fp.N[c] * mag * fp.norm[c];
                        }
                }
                out.any_element_ext = true;
            } else {
                if (uel) continue;                                    // solver_direct.f90:510
                if (d.flag >= 4) continue;
                double tr[3] = {0, 0, 0};
                if (d.flag < 3) for (size_t i = 0; i < d.values.size() && i < 3; ++i) tr[i] = d.values[i];
                else tr[0] = 1.0;
                if (d.flag == 2) {
                    double nr = std::sqrt(tr[0] * tr[0] + tr[1] * tr[1] + tr[2] * tr[2]);
                    for (double& t : tr) t /= nr;
                }
                if (d.flag > 1) for (double& t : tr) t *= hist;
                int nodes[4];
                double fc[4][3];
                face_coords(m, lmn, d.face, nodes, fc);
                for (int kint = 0; kint < 4; ++kint) {
                    FacePoint fp;
                    face_point(fc, kint, fp);
                    const double det = std::sqrt(fp.norm[0] * fp.norm[0] + fp.norm[1] * fp.norm[1] + fp.norm[2] * fp.norm[2]);
                    for (int a = 0; a < 4; ++a)
                        for (int c = 0; c < 3; ++c) {
                            const long ld = loc(3 * (size_t)nodes[a] + c);
                            if (ld < 0) continue;
                            if (d.flag < 3) out.ext[ld] += fp.N[a] * tr[c] * det;
                            else out.ext[ld] -= fp.N[a] * tr[0] * fp.norm[c];
                        }
                }
                out.any_ext = true;
            }
        }
    }

    for (const auto& p : m.prescribed_forces) {
        double fv = p.flag == 1 ? p.value : interpolate_history_table(m.histories[p.history_number - 1], time + dtime);
        auto addf = [&](int n) { const long ld = loc(3 * (size_t)(n - 1) + p.dof - 1); if (ld >= 0) out.ext[ld] += fv; };
        if (p.node_set == 0) addf(p.node_number);
        else for (int n : m.nodesets[p.node_set - 1].items) addf(n);
        out.any_ext = true;
    }

    for (const auto& p : m.prescribed_dofs) {
        double v = p.flag == 1 ? p.value
                 : p.flag == 2 ? interpolate_history_table(m.histories[p.history_number - 1], time + dtime) : 0.0;
        auto one = [&](int n) {
            const long ld = loc(3 * (size_t)(n - 1) + p.dof - 1);
            if (ld < 0) return;
            const int d = (int)ld;
            if (p.flag == 3) {                                        // solver_direct.f90:703-708, :723-728
                double ut3[3] = {0, 0, 0};
                for (int c = 0; c < 3; ++c) { const long lc = loc(3 * (size_t)(n - 1) + c); if (lc >= 0 && dof_total) ut3[c] = dof_total[lc]; }
                bool ignoredof = false;
                user_prescribeddof(m, n, p.dof, m.subroutine_parameters[p.subroutine_parameter_number - 1].values,
                                   &m.coords[3 * (size_t)(n - 1)], ut3, v, ignoredof);
                if (ignoredof) return;
            }
            double target = v;
            // RATE: the value prescribes the increment (solver_direct.f90:706-708); expressed as a target for
            // dof_total + dof_increment so the device forms one kind of correction
            if (p.rate_flag) target = v + (dof_total ? dof_total[d] : 0.0);
            out.fixed_dof.push_back(d);
            out.fixed_value.push_back(target);
            const double ucur = (dof_total ? dof_total[d] : 0.0) + (dof_increment ? dof_increment[d] : 0.0);
            out.fixed_correction.push_back(target - ucur);
        };
        if (p.node_set == 0) one(p.node_number);
        else for (int n : m.nodesets[p.node_set - 1].items) one(n);
    }
}

bool dynamic_load_tables(const Model& m, DynamicLoads& out, const std::vector<int>* g2l, const std::vector<int>* local_elements) {
    out = DynamicLoads();
    const size_t nd = (size_t)m.ndof();
    auto loc = [&](size_t gd) -> long { return g2l ? (long)(*g2l)[gd] : (long)gd; };
    for (const History& h : m.histories) {
        out.history_time.insert(out.history_time.end(), h.t.begin(), h.t.end());
        out.history_value.insert(out.history_value.end(), h.v.begin(), h.v.end());
        out.history_ptr.push_back((int)out.history_time.size());
    }
    const int n_el = local_elements ? (int)local_elements->size() : m.n_elements;
    out.element_density.assign(n_el, 0.0);
    for (int le = 0; le < n_el; ++le) {
        const int e = local_elements ? (*local_elements)[le] : le;
        if (m.element_flag[e] > 99999) {                              // VUEL: amass = amass*props(3) (abaqus_vuel.for:208)
            if (m.element_n_props[e] < 3) { out.error = "a VUEL element needs three properties (E, nu, density)"; return false; }
            out.element_density[le] = m.element_properties[m.element_prop_index[e] + 2];
        } else {
            out.element_density[le] = e < (int)m.element_density.size() ? m.element_density[e] : 0.0;
            if (out.element_density[le] == 0.0) { out.error = "a density has not been defined for element number " + std::to_string(e + 1); return false; }
        }
    }
    std::vector<double> dense(nd);
    auto push_load = [&](int history, double constant) {
        for (size_t d = 0; d < nd; ++d)
            if (dense[d] != 0.0 && loc(d) >= 0) { out.load_dof.push_back((int)loc(d)); out.load_val.push_back(dense[d]); }
        out.load_ptr.push_back((int)out.load_dof.size());
        out.load_history.push_back(history);
        out.load_constant.push_back(constant);
    };
    const bool last_has_history = !m.distributed_loads.empty() && m.distributed_loads.back().history_number > 0;
    for (const auto& d : m.distributed_loads) {
        std::fill(dense.begin(), dense.end(), 0.0);
        const NamedSet& es = m.elementsets[d.element_set - 1];
        int history = 0;
        double constant = 1.0;
        for (int e1 : es.items) {
            const int lmn = e1 - 1;
            int nodes[4];
            double fc[4][3];
            if (m.element_flag[lmn] > 99999) {
                // VUEL, lflags(3) = 3 (:806-905): RHS -= N2(i) adlmag norm(1:3) w; adlmag = history(TIME+DTIME) when the LAST
                // load of the list has a history (:880), else the load's value
                if (!d.uel_format) { out.error = "a native distributed load was applied to a VUEL element"; return false; }
                if (last_has_history) {
                    if (d.history_number <= 0) { out.error = "distributed load without a history while the last load has one (explicit_dynamic_step.f90:880)"; return false; }
                    history = d.history_number;
                } else constant = d.values.empty() ? 0.0 : d.values[0];
                face_coords(m, lmn, d.flag < 0 ? -d.flag : d.flag, nodes, fc);
                for (int kint = 0; kint < 4; ++kint) {
                    FacePoint fp;
                    face_point(fc, kint, fp);
                    for (int i = 0; i < 4; ++i)
                        for (int c = 0; c < 3; ++c) dense[3 * (size_t)nodes[i] + c] -= fp.N[i] * fp.norm[c];
                }
            } else {
                if (d.uel_format) { out.error = "an ABAQUS-format distributed load was applied to a native element"; return false; }
                if (d.flag >= 4) { out.error = "SUBROUTINE distributed loads are not supported in explicit dynamics"; return false; }
                double tr[3] = {0, 0, 0};
                if (d.flag < 3) for (size_t i = 0; i < d.values.size() && i < 3; ++i) tr[i] = d.values[i];
                else tr[0] = 1.0;
                if (d.flag == 2) {
                    const double nr = std::sqrt(tr[0] * tr[0] + tr[1] * tr[1] + tr[2] * tr[2]);
                    for (double& t : tr) t /= nr;
                }
                if (d.flag > 1) history = d.history_number;
                face_coords(m, lmn, d.face, nodes, fc);
                for (int kint = 0; kint < 4; ++kint) {
                    FacePoint fp;
                    face_point(fc, kint, fp);
                    const double det = std::sqrt(fp.norm[0] * fp.norm[0] + fp.norm[1] * fp.norm[1] + fp.norm[2] * fp.norm[2]);
                    for (int a = 0; a < 4; ++a) {
                        if (d.flag < 3) for (int c = 0; c < 3; ++c) dense[3 * (size_t)nodes[a] + c] += fp.N[a] * tr[c] * det;
                        // normal load as traction_boundarycondition_dynamic writes it (traction_boundarycondition.f90:186-192)
                        // with the caller's traction = (value, 0, 0): only the first component receives a force
                        else dense[3 * (size_t)nodes[a]] += fp.N[a] * tr[0] * fp.norm[0];
                    }
                }
            }
        }
        push_load(history, constant);
    }
    for (const auto& p : m.prescribed_forces) {                       // :1029-1063
        if (p.flag == 3) { out.error = "SUBROUTINE prescribed forces are not supported in explicit dynamics"; return false; }
        std::fill(dense.begin(), dense.end(), 0.0);
        auto addf = [&](int n) { dense[3 * (size_t)(n - 1) + p.dof - 1] += 1.0; };
        if (p.node_set == 0) addf(p.node_number);
        else for (int n : m.nodesets[p.node_set - 1].items) addf(n);
        push_load(p.flag == 2 ? p.history_number : 0, p.flag == 1 ? p.value : 1.0);
    }
    std::vector<int> slot(nd, -1);                                    // :1064-1110, later entries overwrite earlier ones
    for (const auto& p : m.prescribed_dofs) {
        if (p.flag == 3) { out.error = "SUBROUTINE prescribed DOFs are not supported in explicit dynamics"; return false; }
        auto one = [&](int n, int mode) {
            const size_t d = 3 * (size_t)(n - 1) + p.dof - 1;
            if (loc(d) < 0) return;
            if (slot[d] < 0) {
                slot[d] = (int)out.pdof_dof.size();
                out.pdof_dof.push_back((int)loc(d)); out.pdof_history.push_back(0); out.pdof_value.push_back(0.0); out.pdof_mode.push_back(0);
            }
            const int k = slot[d];
            out.pdof_history[k] = p.flag == 2 ? p.history_number : 0;
            out.pdof_value[k] = p.flag == 1 ? p.value : 0.0;
            out.pdof_mode[k] = mode;
        };
        if (p.node_set == 0) one(p.node_number, p.rate_flag == 0 ? 0 : 2);     // single node, rate: du = value (:1073-1088)
        else for (int n : m.nodesets[p.node_set - 1].items) one(n, p.rate_flag == 0 ? 0 : 1);
    }
    return true;
}

}  // namespace en234
