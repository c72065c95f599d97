// ORACLE — test infrastructure only (see or_model.h).  Explicit dynamic step (src/explicit_dynamic_step.f90): central
// differences with a row-sum lumped mass matrix; the element loop of assemble_dynamic_force is the residual half of the
// static element loop.  hex8 elements 1001 (el_linelast_3dbasic_dynamic, el_linelast_3Dbasic.f90:140-262; lumped mass
// user_element_lumped_mass, user_element.f90:275-412) and the VUEL of user_codes/src/abaqus_vuel.for ("VU1", density =
// third property).  PARITY PINNING: the reference ships no output of an explicit run: "parity unpinned"; checked against
// the analytic behaviour of a bar (energy balance, wave speed) in tests/.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "or_model.h"

namespace oracle {

static const double SGN[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};

static double hist(const Model& m, int h, double t) {
    const int o = m.history_ptr[h - 1], nh = m.history_ptr[h] - o;
    return interpolate_history_table(&m.history_time[o], &m.history_value[o], nh, t);
}

// assemble_lumped_mass (explicit_dynamic_step.f90:85-392): per element the row sums of the 27-point consistent matrix
// times the density (user_element.f90:380-410 = abaqus_vuel.for:164-208), added node by node in element order
void assemble_lumped_mass(const Model& m, std::vector<double>& lumped_mass) {
    lumped_mass.assign((size_t)m.ndof(), 0.0);
    const double x1[3] = {(double)-0.7745966692f, 0.0, (double)0.7745966692f};
    const double w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        double x[8][3], em[8] = {0, 0, 0, 0, 0, 0, 0, 0};
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
                    for (int a = 0; a < 8; ++a) {
                        const double f1 = 1.0 + SGN[a][0] * xi[0], f2 = 1.0 + SGN[a][1] * xi[1], f3 = 1.0 + SGN[a][2] * xi[2];
                        Nn[a] = f1 * f2 * f3 / 8.0;
                        dN[a][0] = SGN[a][0] * (f2 * f3 / 8.0); dN[a][1] = SGN[a][1] * (f1 * f3 / 8.0); dN[a][2] = SGN[a][2] * (f1 * f2 / 8.0);
                    }
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) { double t = 0; for (int a = 0; a < 8; ++a) t += x[a][r] * dN[a][c]; J[r][c] = t; }
                    const double det = J[0][0] * J[1][1] * J[2][2] - J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] +
                                       J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] - J[0][2] * J[1][1] * J[2][0];
                    double sumN = 0.0;
                    for (int a = 0; a < 8; ++a) sumN += Nn[a];
                    for (int a = 0; a < 8; ++a) em[a] = em[a] + Nn[a] * sumN * det * w;
                }
        const int flag = m.element_flag[lmn];
        const double density = flag > FLAG_UEL_BASE ? m.element_properties[m.element_prop_index[lmn] + 2] : m.element_density[lmn];
        for (int a = 0; a < 8; ++a)
            for (int i = 0; i < 3; ++i) lumped_mass[3 * (size_t)nodes[a] + i] = lumped_mass[3 * (size_t)nodes[a] + i] + em[a] * density;
    }
}

// apply_dynamic_boundaryconditions (explicit_dynamic_step.f90:720-1128): external forces into rforce, prescribed DOFs into
// dof_increment.  Returns non-zero for an option outside the restated subset.
static int apply_dynamic_boundaryconditions(const Model& m, DynamicState& s, double dtime) {
    const int nloads = (int)m.dload_flag.size();
    for (int load = 0; load < nloads; ++load) {
        const int flag = m.dload_flag[load], es = m.dload_elset[load] - 1, ifac = m.dload_face[load];
        for (int k = m.elementset_ptr[es]; k < m.elementset_ptr[es + 1]; ++k) {
            const int lmn = m.elementset_elements[k] - 1;
            double xc[24];
            for (int a = 0; a < 8; ++a)
                for (int i = 0; i < 3; ++i) xc[3 * a + i] = m.coords[3 * (size_t)(m.connectivity[8 * (size_t)lmn + a] - 1) + i];
            if (m.element_flag[lmn] > FLAG_UEL_BASE) {                 // VUEL, lflags(3) = 3 (:806-905)
                double adlmag;
                if (m.dload_history[nloads - 1] > 0) adlmag = hist(m, m.dload_history[load], s.time + dtime);   // :880 tests the LAST load
                else adlmag = m.dload_values[m.dload_val_ptr[load]];
                double R[24];
                vuel_pressure_3d(xc, flag, adlmag, R);
                for (int a = 0; a < 8; ++a) {
                    const int n = m.connectivity[8 * (size_t)lmn + a] - 1;
                    for (int i = 0; i < 3; ++i) s.rforce[3 * (size_t)n + i] = s.rforce[3 * (size_t)n + i] + R[3 * a + i];
                }
                continue;
            }
            if (flag >= 4) return 1;                                   // user_distributed_load_dynamic
            double traction[3] = {0, 0, 0};
            int ntract;
            if (flag < 3) {
                ntract = m.dload_val_ptr[load + 1] - m.dload_val_ptr[load];
                for (int i = 0; i < ntract && i < 3; ++i) traction[i] = m.dload_values[m.dload_val_ptr[load] + i];
            } else { ntract = 1; traction[0] = 1.0; }
            if (flag == 2) {
                double nr = 0.0;
                for (int i = 0; i < ntract; ++i) nr += traction[i] * traction[i];
                nr = std::sqrt(nr);
                for (int i = 0; i < ntract; ++i) traction[i] = traction[i] / nr;
            }
            if (flag > 1) {
                const double v = hist(m, m.dload_history[load], s.time + dtime);
                for (double& t : traction) t = t * v;
            }
            int list[4];
            facenodes_hex8(ifac, list);
            double fc[12], R12[12];
            for (int j = 0; j < 4; ++j)
                for (int i = 0; i < 3; ++i) fc[3 * j + i] = xc[3 * (list[j] - 1) + i];
            traction_boundarycondition_dynamic_3d(flag, fc, traction, ntract, R12);
            for (int j = 0; j < 4; ++j) {
                const int n = m.connectivity[8 * (size_t)lmn + list[j] - 1] - 1;
                for (int i = 0; i < 3; ++i) s.rforce[3 * (size_t)n + i] = s.rforce[3 * (size_t)n + i] + R12[3 * j + i];
            }
        }
    }
    for (size_t k = 0; k < m.pforce_flag.size(); ++k) {               // :1029-1063
        double fv = 0.0;
        if (m.pforce_flag[k] == 1) fv = m.pforce_value[k];
        else if (m.pforce_flag[k] == 2) fv = hist(m, m.pforce_history[k], s.time + dtime);
        else return 1;                                                 // user_prescribedforce
        const int idof = m.pforce_dof[k] - 1;
        if (m.pforce_nodeset[k] == 0) s.rforce[3 * (size_t)(m.pforce_node[k] - 1) + idof] += fv;
        else {
            const int ns = m.pforce_nodeset[k] - 1;
            for (int i = m.nodeset_ptr[ns]; i < m.nodeset_ptr[ns + 1]; ++i) s.rforce[3 * (size_t)(m.nodeset_nodes[i] - 1) + idof] += fv;
        }
    }
    for (size_t k = 0; k < m.pdof_flag.size(); ++k) {                 // :1064-1110
        double v = 0.0;
        if (m.pdof_flag[k] == 1) v = m.pdof_value[k];
        else if (m.pdof_flag[k] == 2) v = hist(m, m.pdof_history[k], s.time + dtime);
        else return 1;                                                 // user_prescribeddof
        const int idof = m.pdof_dof[k] - 1;
        if (m.pdof_nodeset[k] == 0) {
            // single node (:1073-1088): the increment is value - ucur, with ucur = 0 for a rate (no DTIME factor there)
            const size_t d = 3 * (size_t)(m.pdof_node[k] - 1) + idof;
            const double ucur = m.pdof_rate[k] == 0 ? s.dof_total[d] : 0.0;
            s.dof_increment[d] = v - ucur;
        } else {
            const int ns = m.pdof_nodeset[k] - 1;
            for (int i = m.nodeset_ptr[ns]; i < m.nodeset_ptr[ns + 1]; ++i) {
                const size_t d = 3 * (size_t)(m.nodeset_nodes[i] - 1) + idof;
                if (m.pdof_rate[k] == 0) s.dof_increment[d] = v - s.dof_total[d];
                else s.dof_increment[d] = dtime * v;                  // :1103-1106
            }
        }
    }
    return 0;
}

// assemble_dynamic_force (explicit_dynamic_step.f90:394-718)
static int assemble_dynamic_force(const Model& m, DynamicState& s) {
    std::vector<double> K(576);
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        double xc[24], du[24], ut[24], R[24];
        int nodes[8];
        for (int a = 0; a < 8; ++a) {
            nodes[a] = m.connectivity[8 * (size_t)lmn + a] - 1;
            for (int i = 0; i < 3; ++i) {
                xc[3 * a + i] = m.coords[3 * (size_t)nodes[a] + i];
                du[3 * a + i] = s.dof_increment[3 * (size_t)nodes[a] + i];
                ut[3 * a + i] = s.dof_total[3 * (size_t)nodes[a] + i];
            }
        }
        const double* props = m.element_properties.data() + m.element_prop_index[lmn];
        const int flag = m.element_flag[lmn];
        if (flag == ID_LINELAST) el_linelast_3dbasic(xc, du, ut, props, K.data(), R);    // residual part = el_linelast_3dbasic_dynamic
        else if (flag == FLAG_UEL_BASE + 1) {                          // VUEL lflags(3) = 2: U = dof_total + dof_increment (:560-566)
            double U[24], pn;
            for (int q = 0; q < 24; ++q) U[q] = du[q] + ut[q];
            uel_linelastic_3d(xc, U, props, 0, 0, nullptr, nullptr, K.data(), R, nullptr, nullptr, &pn);
        } else return 1;
        for (int a = 0; a < 8; ++a)
            for (int i = 0; i < 3; ++i) s.rforce[3 * (size_t)nodes[a] + i] = s.rforce[3 * (size_t)nodes[a] + i] + R[3 * a + i];
    }
    return 0;
}

int explicit_dynamic_steps(const Model& m, DynamicState& s, int n_steps) {
    const size_t nd = (size_t)m.ndof();
    const double DTIME = m.dynamic_timestep;
    if (s.steps_done == 0 && s.velocity.empty()) {                     // explicit_dynamic_step.f90:24-33
        s.dof_total.resize(nd, 0.0);
        s.dof_increment.resize(nd, 0.0);
        s.velocity = s.dof_increment;
        s.acceleration.assign(nd, 0.0);
        s.time = 0.0;
    }
    s.rforce.assign(nd, 0.0);
    for (int istep = 0; istep < n_steps; ++istep) {
        for (size_t d = 0; d < nd; ++d) s.dof_increment[d] = DTIME * (s.velocity[d] + 0.5 * DTIME * s.acceleration[d]);
        std::fill(s.rforce.begin(), s.rforce.end(), 0.0);
        if (s.lumped_mass.empty()) assemble_lumped_mass(m, s.lumped_mass);   // the reference recomputes it every step; same values
        if (int rc = apply_dynamic_boundaryconditions(m, s, DTIME)) return rc;
        if (int rc = assemble_dynamic_force(m, s)) return rc;
        for (size_t d = 0; d < nd; ++d) {
            s.acceleration[d] = s.rforce[d] / s.lumped_mass[d];
            s.velocity[d] = s.velocity[d] + DTIME * s.acceleration[d];
            s.dof_total[d] = s.dof_total[d] + s.dof_increment[d];
        }
        s.time = s.time + DTIME;
        s.steps_done++;
    }
    return 0;
}

}  // namespace oracle
