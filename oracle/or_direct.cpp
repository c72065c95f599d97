// ORACLE — test infrastructure only (see or_model.h).  Skyline ("profile") direct solver path:
// restates src/solver_direct.f90 (assembly :5-439, BCs :444-738, fixdof :742-792, solve :916-973,
// back-substitution :976-1032, LDU factorisation :1036-1192, profile :1196-1393) with 0-based arrays.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "or_model.h"

namespace oracle {

// Column j of the upper triangle holds rows j-h..j-1 at aupp[jp[j-1] .. jp[j]-1] (h = jp[j]-jp[j-1]).
static inline long long colstart(const Skyline& s, int j) { return j == 0 ? 0 : s.jp[j - 1]; }
static inline int colheight(const Skyline& s, int j) { return (int)(s.jp[j] - colstart(s, j)); }

// compute_profile (solver_direct.f90:1252-1360): equations numbered in node_order, 3 per node;
// column height = max over elements of (eq - smallest eq on the element).
void compute_profile(const Model& m, const std::vector<int>& node_order, Skyline& s, StepLog* log) {
    // constraints: one extra equation each, ieqs[length_dofs + nc] (:1252-1266).  The reference orders them with the nodes
    // (the Gibbs-Poole-Stockmeyer pass treats a constraint as a node); here node_order may list n_nodes + n_constraints
    // entries (an entry >= n_nodes is constraint entry - n_nodes), otherwise the constraint equations follow all DOFs.
    const int ncon = m.n_constraints();
    const int neq = m.ndof() + ncon;
    s.neq = neq;
    s.ieqs.assign(neq, 0);
    s.lagrange_multipliers.assign(ncon, 0.0);
    int e = 0;
    const bool with_con = (int)node_order.size() == m.n_nodes + ncon;
    for (int i = 0; i < (with_con ? m.n_nodes + ncon : m.n_nodes); ++i) {
        int n = node_order[i];
        if (n < m.n_nodes) for (int j = 0; j < 3; ++j) s.ieqs[3 * n + j] = e++;
        else s.ieqs[m.ndof() + n - m.n_nodes] = e++;
    }
    if (!with_con) for (int nc = 0; nc < ncon; ++nc) s.ieqs[m.ndof() + nc] = e++;
    std::vector<long long> h(neq, 0);
    const int npe = m.nodes_per_element;
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        int mineq = neq - 1;
        for (int i = 0; i < npe; ++i) {
            int n = m.connectivity[(size_t)npe * lmn + i] - 1;
            for (int j = 0; j < 3; ++j) mineq = std::min(mineq, s.ieqs[3 * n + j]);
        }
        for (int i = 0; i < npe; ++i) {
            int n = m.connectivity[(size_t)npe * lmn + i] - 1;
            for (int j = 0; j < 3; ++j) {
                int q = s.ieqs[3 * n + j];
                h[q] = std::max<long long>(h[q], q - mineq);
            }
        }
    }
    for (int nc = 0; nc < ncon; ++nc) {                       // two-node ties :1296-1317
        const int q[3] = {s.ieqs[3 * (m.con_node1[nc] - 1) + m.con_dof1[nc] - 1], s.ieqs[3 * (m.con_node2[nc] - 1) + m.con_dof2[nc] - 1],
                          s.ieqs[m.ndof() + nc]};
        const int mineq = std::min(q[0], std::min(q[1], q[2]));
        for (int k = 0; k < 3; ++k) h[q[k]] = std::max<long long>(h[q[k]], q[k] - mineq);
    }
    // statistics exactly as :1339-1351 (first column excluded from max/mean/rms sums, divided by neq)
    long long bwmax = 0;
    double bwmn = 0.0, bwrms = 0.0;
    s.jp.assign(neq, 0);
    h[0] = 0;
    for (int n = 1; n < neq; ++n) {
        bwmax = std::max(bwmax, h[n]);
        bwmn += (double)h[n];
        bwrms += (double)h[n] * (double)h[n];
    }
    long long acc = 0;
    for (int n = 0; n < neq; ++n) { acc += h[n]; s.jp[n] = acc; }
    if (log) {
        log->profile_neq = neq;
        log->profile_bwmax = bwmax;
        log->profile_bwmean = (long long)(bwmn / (double)neq + 0.5);
        log->profile_bwrms = (long long)(std::sqrt(bwrms / (double)neq) + 0.5);
        log->profile_size = s.jp[neq - 1];
    }
    s.aupp.assign((size_t)s.jp[neq - 1], 0.0);
    s.unsym = m.unsymmetric != 0;
    if (s.unsym) s.alow.assign((size_t)s.jp[neq - 1], 0.0); else s.alow.clear();
    s.diag.assign(neq, 0.0);
    s.rhs.assign(neq, 0.0);
}

// assemble_direct_stiffness (solver_direct.f90:95-423)
int assemble_direct_stiffness(const Model& m, Skyline& s, const double* dof_total, const double* dof_increment,
                              double time, double dtime, std::vector<double>* svars, AsmOut& out) {
    std::fill(s.aupp.begin(), s.aupp.end(), 0.0);
    std::fill(s.alow.begin(), s.alow.end(), 0.0);
    std::fill(s.rhs.begin(), s.rhs.end(), 0.0);
    std::fill(s.diag.begin(), s.diag.end(), 0.0);
    const int npe = m.nodes_per_element, nd = 3 * npe;
    std::vector<double> Kbuf((size_t)nd * nd), Rbuf(nd);
    double *K = Kbuf.data(), *R = Rbuf.data();
    long long sv_off = 0;
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        double* sv = nullptr;
        if (svars && m.element_n_states[lmn] > 0) sv = svars->data() + sv_off;
        sv_off += m.element_n_states[lmn];
        if (element_static(m, lmn, dof_total, dof_increment, time, dtime, K, R, sv)) return 1;
        // scatter :264-287
        for (int i1 = 0; i1 < npe; ++i1) {
            int node1 = m.connectivity[(size_t)npe * lmn + i1] - 1;
            for (int j1 = 0; j1 < 3; ++j1) {
                int irow = s.ieqs[3 * node1 + j1], nsr = 3 * i1 + j1;
                s.rhs[irow] += R[nsr];
                for (int i2 = 0; i2 < npe; ++i2) {
                    int node2 = m.connectivity[(size_t)npe * lmn + i2] - 1;
                    for (int j2 = 0; j2 < 3; ++j2) {
                        int icol = s.ieqs[3 * node2 + j2], nsc = 3 * i2 + j2;
                        if (icol == irow) s.diag[irow] += K[nsr + nd * nsc];
                        else if (icol > irow) {
                            const size_t nn = (size_t)(s.jp[icol] - (icol - irow));
                            s.aupp[nn] += K[nsr + nd * nsc];
                            if (s.unsym) s.alow[nn] += K[nsc + nd * nsr];          // :282
                        }
                    }
                }
            }
        }
    }
    // two-node tie constraints :292-328 (the multiplier's diagonal is a penalty-scaled compliance 1e-12 |diag| / neq)
    if (m.n_constraints() > 0) {
        double dd = 0.0;
        for (int n = 0; n < s.neq; ++n) dd += s.diag[n] * s.diag[n];
        const double diagnorm = std::sqrt(dd) / s.neq;
        auto upper = [&](int a, int b) -> size_t { return a > b ? (size_t)(s.jp[a] - (a - b)) : (size_t)(s.jp[b] - (b - a)); };
        for (int nc = 0; nc < m.n_constraints(); ++nc) {
            const int icol = s.ieqs[m.ndof() + nc];
            const int iof1 = 3 * (m.con_node1[nc] - 1) + m.con_dof1[nc] - 1, iof2 = 3 * (m.con_node2[nc] - 1) + m.con_dof2[nc] - 1;
            const double lam = s.lagrange_multipliers[nc];
            int irow = s.ieqs[iof1];
            s.rhs[irow] -= lam;
            s.aupp[upper(icol, irow)] += 1.0;
            if (s.unsym) s.alow[upper(icol, irow)] += 1.0;
            irow = s.ieqs[iof2];
            s.rhs[irow] += lam;
            s.aupp[upper(icol, irow)] -= 1.0;
            if (s.unsym) s.alow[upper(icol, irow)] -= 1.0;
            s.rhs[icol] = dof_total[iof2] + dof_increment[iof2] - dof_total[iof1] - dof_increment[iof1] - lam * 1.e-12 * diagnorm;
            s.diag[icol] = 1.e-12 * diagnorm;
        }
    }
    // NaN scan :402-415
    for (int n = 0; n < s.neq; ++n)
        if (s.rhs[n] == s.rhs[n] + 1.0 || std::isnan(s.rhs[n])) return 1;
    // rforce = -rhs, nodal_force_norm :417-423
    out.rforce.assign(m.ndof(), 0.0);
    double nn = 0.0;
    for (int d = 0; d < m.ndof(); ++d) out.rforce[d] = -s.rhs[s.ieqs[d]];
    for (int n = 0; n < s.neq; ++n) nn += s.rhs[n] * s.rhs[n];
    out.nodal_force_norm = std::sqrt(nn);
    return 0;
}

// fixdof_direct (solver_direct.f90:742-792), symmetric case
static void fixdof_direct(Skyline& s, int n, double dofvalue) {
    for (int ncol = n + 1; ncol < s.neq; ++ncol) {
        int idis = ncol - n;
        if (idis <= colheight(s, ncol)) {
            size_t nn = (size_t)(s.jp[ncol] - idis);
            if (!s.unsym) s.rhs[ncol] -= s.aupp[nn] * dofvalue;          // :767
            s.aupp[nn] = 0.0;
        }
    }
    if (s.unsym) {                                                       // :772-777: only row n is replaced
        for (long long i = colstart(s, n); i < s.jp[n]; ++i) s.alow[(size_t)i] = 0.0;
    } else if (n > 0) {
        int mod = n;
        for (long long i = s.jp[n] - 1; i >= colstart(s, n); --i) {
            --mod;
            s.rhs[mod] -= s.aupp[(size_t)i] * dofvalue;
            s.aupp[(size_t)i] = 0.0;
        }
    }
    s.diag[n] = 1.0;
    s.rhs[n] = dofvalue;
}

// user_prescribeddof as shipped (user_codes/src/user_prescribeddof.f90:48-103): a bar is stretched along tensile_dir over the
// first n_extension_steps steps and then rotated rigidly about rotation_axis in n_rotation_steps steps.  Returns false when
// the routine sets ignoredof.  Parameter list: :48-58; nodes on the base / end follow from entry 12 on.
static bool oracle_user_prescribeddof(const Model& m, int node_number, int idof, int k, const double* dof_total, double* dofvalue) {
    const int sp = k < (int)m.pdof_subparam.size() ? m.pdof_subparam[k] : 0;
    if (sp < 1) return false;
    const double* par = &m.subparam_values[m.subparam_ptr[sp - 1]];
    const int npar = m.subparam_ptr[sp] - m.subparam_ptr[sp - 1];
    if (npar < 11) return false;
    const int tensile_dir = (int)par[0], n_ext = (int)par[2], axis_no = (int)par[4], n_rot = (int)par[5];
    const int n_base = (int)par[6], n_end = (int)par[7];
    bool ignoredof = true;
    const int step = m.current_step_number;
    if (step <= n_ext) {
        for (int i = 12; i <= 12 + n_base - 1 && i <= npar; ++i)
            if (node_number == (int)par[i - 1] && idof == tensile_dir) { *dofvalue = 0.0; ignoredof = false; }
        for (int i = 12 + n_base; i <= 12 + n_base + n_end - 1 && i <= npar; ++i)
            if (node_number == (int)par[i - 1] && idof == tensile_dir) { *dofvalue = par[1] * step / n_ext; ignoredof = false; }
        if (node_number == (int)par[8]) { *dofvalue = 0.0; ignoredof = false; }
        if (node_number == (int)par[9] && idof == (int)par[10]) { *dofvalue = 0.0; ignoredof = false; }
    } else if (step <= n_ext + n_rot) {
        double ref[3], cur[3], ax[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < 3; ++i) { ref[i] = m.coords[3 * (size_t)(node_number - 1) + i]; cur[i] = ref[i] + dof_total[3 * (size_t)(node_number - 1) + i]; }
        const double angle = 3.141592653589793238462643383279502884197 * par[3] / (n_rot * 180.0);
        ax[axis_no - 1] = 1.0;
        const double dot = ax[0] * cur[0] + ax[1] * cur[1] + ax[2] * cur[2];
        const double cross[3] = {ax[1] * cur[2] - ax[2] * cur[1], cur[0] * ax[2] - ax[0] * cur[2], ax[0] * cur[1] - ax[1] * cur[0]};
        const double newc = std::cos(angle) * cur[idof - 1] + (1.0 - std::cos(angle)) * dot * ax[idof - 1] + std::sin(angle) * cross[idof - 1];
        *dofvalue = newc - ref[idof - 1];
        ignoredof = false;
    } else { *dofvalue = 0.0; ignoredof = false; }
    return !ignoredof;
}

static double history_value(const Model& m, int h, double t) {
    int o = m.history_ptr[h - 1], nh = m.history_ptr[h] - o;
    return interpolate_history_table(&m.history_time[o], &m.history_value[o], nh, t);
}

// Shared by both solver paths: native-element tractions (solver_direct.f90:501-568) and prescribed
// nodal forces (:646-682) added to rhs through `eq`; returns the list of (dof, correction) pairs of
// the prescribed-DOF block (:685-734) in application order.
static void bc_loads_and_corrections(const Model& m, const double* dof_total, const double* dof_increment,
                                     double time, double dtime, bool nodal_forces, const int* ieqs,
                                     std::vector<double>& rhs, std::vector<std::pair<int, double>>& fix) {
    const int nloads = (int)m.dload_flag.size();
    for (int load = 0; load < nloads; ++load) {
        int flag = m.dload_flag[load], es = m.dload_elset[load] - 1, ifac = m.dload_face[load];
        for (int k = m.elementset_ptr[es]; k < m.elementset_ptr[es + 1]; ++k) {
            int lmn = m.elementset_elements[k] - 1;
            if (m.element_flag[lmn] > FLAG_UEL_BASE) continue;    // :510 skip ABAQUS UEL
            if (flag >= 4) continue;                               // user-subroutine loads: out of scope
            double traction[3] = {0, 0, 0};
            int ntract;
            if (flag < 3) {
                ntract = m.dload_val_ptr[load + 1] - m.dload_val_ptr[load];
                for (int i = 0; i < ntract && i < 3; ++i) traction[i] = m.dload_values[m.dload_val_ptr[load] + i];
            } else { ntract = 1; traction[0] = 1.0; }
            if (flag == 2) {
                double nr = std::sqrt(traction[0] * traction[0] + traction[1] * traction[1] + traction[2] * traction[2]);
                for (double& t : traction) t = t / nr;
            }
            if (flag > 1) {
                double v = history_value(m, m.dload_history[load], time + dtime);
                for (double& t : traction) t = t * v;
            }
            int list[4];
            facenodes_hex8(ifac, list);
            double fc[12], R12[12];
            for (int j = 0; j < 4; ++j) {
                int n = m.connectivity[8 * lmn + list[j] - 1] - 1;
                for (int i = 0; i < 3; ++i) fc[3 * j + i] = m.coords[3 * n + i];
            }
            traction_boundarycondition_static_3d(flag, fc, traction, ntract, R12);
            for (int j = 0; j < 4; ++j) {
                int n = m.connectivity[8 * lmn + list[j] - 1] - 1;
                for (int i = 0; i < 3; ++i) rhs[ieqs[3 * n + i]] += R12[3 * j + i];
            }
        }
    }
    if (nodal_forces) {
        for (size_t k = 0; k < m.pforce_flag.size(); ++k) {
            double fv = 0.0;
            if (m.pforce_flag[k] == 1) fv = m.pforce_value[k];
            else if (m.pforce_flag[k] == 2) fv = history_value(m, m.pforce_history[k], time + dtime);
            int idof = m.pforce_dof[k] - 1;
            if (m.pforce_nodeset[k] == 0) rhs[ieqs[3 * (m.pforce_node[k] - 1) + idof]] += fv;
            else {
                int ns = m.pforce_nodeset[k] - 1;
                for (int i = m.nodeset_ptr[ns]; i < m.nodeset_ptr[ns + 1]; ++i)
                    rhs[ieqs[3 * (m.nodeset_nodes[i] - 1) + idof]] += fv;
            }
        }
    }
    for (size_t k = 0; k < m.pdof_flag.size(); ++k) {
        double v = 0.0;
        if (m.pdof_flag[k] == 1) v = m.pdof_value[k];
        else if (m.pdof_flag[k] == 2) v = history_value(m, m.pdof_history[k], time + dtime);
        int idof = m.pdof_dof[k] - 1;
        auto one = [&](int n) {
            int d = 3 * n + idof;
            double ucur = m.pdof_rate[k] == 0 ? dof_increment[d] + dof_total[d] : dof_increment[d];
            if (m.pdof_flag[k] == 3) {                              // solver_direct.f90:703-708 / :723-728
                double value = 0.0;
                if (!oracle_user_prescribeddof(m, n + 1, idof + 1, (int)k, dof_total, &value)) return;    // ignoredof
                fix.emplace_back(d, value - ucur);
                return;
            }
            fix.emplace_back(d, v - ucur);
        };
        if (m.pdof_nodeset[k] == 0) one(m.pdof_node[k] - 1);
        else {
            int ns = m.pdof_nodeset[k] - 1;
            for (int i = m.nodeset_ptr[ns]; i < m.nodeset_ptr[ns + 1]; ++i) one(m.nodeset_nodes[i] - 1);
        }
    }
}
// exported for or_cg.cpp
void oracle_bc_loads_and_corrections(const Model& m, const double* dof_total, const double* dof_increment,
                                     double time, double dtime, bool nodal_forces, const int* ieqs,
                                     std::vector<double>& rhs, std::vector<std::pair<int, double>>& fix) {
    bc_loads_and_corrections(m, dof_total, dof_increment, time, dtime, nodal_forces, ieqs, rhs, fix);
}

// apply_direct_boundaryconditions (solver_direct.f90:444-738)
void apply_direct_boundaryconditions(const Model& m, Skyline& s, const double* dof_total,
                                     const double* dof_increment, double time, double dtime, AsmOut& out) {
    std::vector<std::pair<int, double>> fix;
    bc_loads_and_corrections(m, dof_total, dof_increment, time, dtime, true, s.ieqs.data(), s.rhs, fix);
    for (auto& f : fix) fixdof_direct(s, s.ieqs[f.first], f.second);
    double nn = 0.0;
    for (int n = 0; n < s.neq; ++n) nn += s.rhs[n] * s.rhs[n];
    out.unbalanced_force_norm = std::sqrt(nn);                       // :736
}

// LU_decomposition + dredu (solver_direct.f90:1036-1192), symmetric case (alow aliases aupp):
// Crout reduction column by column; the scaled column (U/d) replaces aupp, diag holds reciprocals.
static void lu_decomposition(Skyline& s, StepLog* log) {
    const double tol = 0.5e-7;
    const int neq = s.neq;
    double dimx = 0.0, dfig = 0.0, dimn = 0.0;
    for (int i = 0; i < neq; ++i) dimn = std::max(dimn, std::fabs(s.diag[i]));
    double* a = s.aupp.data();
    double* l = s.unsym ? s.alow.data() : s.aupp.data();     // symmetric: alow aliases aupp (solve_direct :956-957)
    for (int j = 0; j < neq; ++j) {
        const long long c0 = colstart(s, j);
        const int h = colheight(s, j);
        const int is = j - h;                      // first stored row of column j
        double daval = 0.0;
        if (h > 1) {
            if (s.diag[j] == 0.0)
                for (int k = 0; k < h; ++k) daval += std::fabs(a[c0 + k]);   // :1100 (sum over jr..jr+jh)
            for (int i = is + 1; i < j; ++i) {     // reduce entry (i,j) :1101-1111
                const int hi = colheight(s, i);
                const int ih = std::min(hi, i - is);
                if (ih > 0) {
                    const long long oj = c0 + (i - is) - ih;     // rows i-ih..i-1 of column j / row j
                    const long long oi = s.jp[i] - ih;            // rows i-ih..i-1 of column i / row i
                    double dot = 0.0;
                    for (int k = 0; k < ih; ++k) dot += a[oj + k] * l[oi + k];
                    if (s.unsym) {                                 // :1109
                        double dl = 0.0;
                        for (int k = 0; k < ih; ++k) dl += l[oj + k] * a[oi + k];
                        a[c0 + (i - is)] -= dot;
                        l[c0 + (i - is)] -= dl;
                    } else a[c0 + (i - is)] -= dot;
                }
            }
        }
        const double dd = s.diag[j];
        if (h >= 1) {                              // dredu :1164-1192
            double dj = s.diag[j];
            for (int k = 0; k < h; ++k) {
                double ud = a[c0 + k] * s.diag[is + k];
                dj = dj - l[c0 + k] * ud;
                a[c0 + k] = ud;
            }
            if (s.unsym) for (int k = 0; k < h; ++k) l[c0 + k] = l[c0 + k] * s.diag[is + k];   // :1186-1188
            s.diag[j] = dj;
        }
        if (s.diag[j] != 0.0) {                    // :1129-1137
            dimx = std::max(dimx, std::fabs(s.diag[j]));
            dimn = std::min(dimn, std::fabs(s.diag[j]));
            dfig = std::max(dfig, std::fabs(dd / s.diag[j]));
            s.diag[j] = 1.0 / s.diag[j];
        } else {
            s.diag[j] = 1.0 / (tol * dimn);
        }
        (void)daval;
    }
    if (log) {
        log->lu_dimx.push_back(dimx);
        log->lu_dimn.push_back(dimn);
        log->lu_ifig.push_back((int)(std::log10(dfig) + 0.6));
    }
}

// The consistent field-projection mass matrix of print_state (src/printing.f90:1214-1383, assemble_projection_mass_matrix,
// hex8 branch: 27-point rule with the reference's literals, Element_Utilities.f90:651-656) in node-level skyline storage
// with the node numbering given, factorised by the same LDU routine (printing.f90:475-488).  Only the conditioning line
// the reference prints for it is of interest here: the golden log carries it once per printed step.
void projection_mass_lu(const Model& m, const std::vector<int>& node_order, StepLog* log) {
    Skyline s;
    projection_mass_factor(m, node_order, s, log);
}
// the factorised matrix itself (print_state then back-substitutes one right-hand side per field variable, :480-488)
void projection_mass_factor(const Model& m, const std::vector<int>& node_order, Skyline& s, StepLog* log) {
    const int N = m.n_nodes;
    std::vector<int> num(N, 0);                       // node -> equation (0-based)
    for (int i = 0; i < N; ++i) num[node_order[i]] = i;
    s = Skyline();
    s.neq = N;
    std::vector<long long> h(N, 0);
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        int mineq = N - 1;
        for (int i = 0; i < 8; ++i) mineq = std::min(mineq, num[m.connectivity[8 * (size_t)lmn + i] - 1]);
        for (int i = 0; i < 8; ++i) {
            const int q = num[m.connectivity[8 * (size_t)lmn + i] - 1];
            h[q] = std::max<long long>(h[q], q - mineq);
        }
    }
    h[0] = 0;
    s.jp.assign(N, 0);
    long long acc = 0;
    for (int n = 0; n < N; ++n) { acc += h[n]; s.jp[n] = acc; }
    s.aupp.assign((size_t)acc, 0.0);
    s.diag.assign(N, 0.0);
    s.unsym = false;
    const double x1[3] = {(double)-0.7745966692f, 0.0, (double)0.7745966692f};
    const double w1[3] = {0.5555555555, 0.888888888, 0.55555555555};
    static const double SGN[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        double x[8][3], em[8][8];
        int nodes[8];
        for (int a = 0; a < 8; ++a) {
            nodes[a] = m.connectivity[8 * (size_t)lmn + a] - 1;
            for (int i = 0; i < 3; ++i) x[a][i] = m.coords[3 * (size_t)nodes[a] + i];
        }
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) em[a][b] = 0.0;
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) {
                    const double xi[3] = {x1[i], x1[j], x1[k]}, w = w1[i] * w1[j] * w1[k];
                    double Nn[8], dN[8][3], J[3][3];
                    for (int a = 0; a < 8; ++a) {                    // Element_Utilities.f90:908-940
                        const double f1 = 1.0 + SGN[a][0] * xi[0], f2 = 1.0 + SGN[a][1] * xi[1], f3 = 1.0 + SGN[a][2] * xi[2];
                        Nn[a] = f1 * f2 * f3 / 8.0;
                        dN[a][0] = SGN[a][0] * (f2 * f3 / 8.0); dN[a][1] = SGN[a][1] * (f1 * f3 / 8.0); dN[a][2] = SGN[a][2] * (f1 * f2 / 8.0);
                    }
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) { double t = 0; for (int a = 0; a < 8; ++a) t += x[a][r] * dN[a][c]; J[r][c] = t; }
                    const double det = J[0][0] * J[1][1] * J[2][2] - J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] +
                                       J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] - J[0][2] * J[1][1] * J[2][0];
                    for (int b = 0; b < 8; ++b)
                        for (int a = 0; a < 8; ++a) em[a][b] = em[a][b] + Nn[b] * Nn[a] * det * w;
                }
        for (int j = 0; j < 8; ++j) {
            const int irow = num[nodes[j]];
            for (int i = 0; i < 8; ++i) {
                const int icol = num[nodes[i]];
                if (icol == irow) s.diag[irow] += em[i][i];
                else if (icol > irow) s.aupp[(size_t)(s.jp[icol] - 1 - (icol - 1 - irow))] += em[j][i];
            }
        }
    }
    lu_decomposition(s, log);
}

// backsubstitution (solver_direct.f90:976-1032)
static void backsubstitution(Skyline& s) {
    const int neq = s.neq;
    double* a = s.aupp.data();
    const double* l = s.unsym ? s.alow.data() : s.aupp.data();
    double* rhs = s.rhs.data();
    int is = 0;
    while (is < neq && rhs[is] == 0.0) ++is;
    if (is == neq) { std::fill(s.rhs.begin(), s.rhs.end(), 0.0); return; }
    for (int j = is + 1; j < neq; ++j) {
        const int h = colheight(s, j);
        if (h > 0) {
            const double* cj = l + colstart(s, j);
            double dot = 0.0;
            for (int k = 0; k < h; ++k) dot += cj[k] * rhs[j - h + k];
            rhs[j] -= dot;
        }
    }
    for (int j = is; j < neq; ++j) rhs[j] = rhs[j] * s.diag[j];
    for (int j = neq - 1; j >= 1; --j) {
        const int h = colheight(s, j);
        if (h > 0) {
            const double* cj = a + colstart(s, j);
            const double rj = rhs[j];
            for (int k = 0; k < h; ++k) rhs[j - h + k] -= rj * cj[k];
        }
    }
}

void skyline_backsubstitution(Skyline& s) { backsubstitution(s); }

// solve_direct (solver_direct.f90:916-973)
void solve_direct(const Model& m, Skyline& s, double* dof_increment, AsmOut& out, StepLog* log) {
    lu_decomposition(s, log);
    backsubstitution(s);
    for (int d = 0; d < m.ndof(); ++d) dof_increment[d] += s.rhs[s.ieqs[d]];
    for (int nc = 0; nc < m.n_constraints(); ++nc) s.lagrange_multipliers[nc] += s.rhs[s.ieqs[m.ndof() + nc]];   // :967-969
    double nn = 0.0;
    for (int n = 0; n < s.neq; ++n) nn += s.rhs[n] * s.rhs[n];
    out.correction_norm = std::sqrt(nn);
}

}  // namespace oracle
