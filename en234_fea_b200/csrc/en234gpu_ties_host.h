// Host side of the tie-constraint layer (kernels and the method: en234gpu_ties.cuh).  Included by en234gpu.cu.
#pragma once

static TieTables tie_tables(en234gpu_ctx* c) {
    Ties& t = c->ties;
    TieTables T;
    T.counts = t.counts.p; T.g_root = t.g_root.p; T.g_ptr = t.g_ptr.p; T.g_fixed = t.g_fixed.p;
    T.m_dof = t.m_dof.p; T.m_root = t.m_root.p; T.m_fixed = t.m_fixed.p;
    return T;
}

// Groups of tied DOFs for a prescribed set — pure host logic (no device): a group with a prescribed member follows that member
// (its root), every other group is reduced to its smallest DOF.  Also the order in which the multiplier corrections are
// recovered after a solve (leaves of every constraint tree first; a prescribed DOF's row carries no information).
struct TiePlan {
    std::vector<int> tdof, idx1, idx2;                 // touched DOFs (sorted), index of each constraint's two DOFs in that list
    std::vector<int> g_root, g_ptr{0}, m_dof, m_root;  // groups: root DOF, members (root excluded, ascending DOF order)
    std::vector<uint8_t> g_fixed, m_fixed;
    std::vector<int> peel_tie, peel_leaf;              // recovery order: constraint, index (in tdof) of its leaf end
    std::string error;
};

static bool tie_plan(int n, const int* d1, const int* d2, const std::vector<int>& fixdof, TiePlan& P) {
    P = TiePlan();
    P.tdof.assign(d1, d1 + n);
    P.tdof.insert(P.tdof.end(), d2, d2 + n);
    std::sort(P.tdof.begin(), P.tdof.end());
    P.tdof.erase(std::unique(P.tdof.begin(), P.tdof.end()), P.tdof.end());
    const int nt = (int)P.tdof.size();
    auto index_of = [&](int d) { return (int)(std::lower_bound(P.tdof.begin(), P.tdof.end(), d) - P.tdof.begin()); };
    P.idx1.resize(n); P.idx2.resize(n);
    for (int k = 0; k < n; ++k) { P.idx1[k] = index_of(d1[k]); P.idx2[k] = index_of(d2[k]); }
    std::vector<int> parent(nt);
    for (int i = 0; i < nt; ++i) parent[i] = i;
    auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
    for (int k = 0; k < n; ++k) {
        const int a = find(P.idx1[k]), b = find(P.idx2[k]);
        if (a == b) { P.error = "tie constraints: constraint " + std::to_string(k + 1) + " closes a loop (redundant constraint)"; return false; }
        parent[std::max(a, b)] = std::min(a, b);       // the smallest touched DOF is the representative
    }
    std::vector<uint8_t> fixed(nt, 0);
    for (int d : fixdof) {
        auto it = std::lower_bound(P.tdof.begin(), P.tdof.end(), d);
        if (it != P.tdof.end() && *it == d) fixed[it - P.tdof.begin()] = 1;
    }
    std::vector<int> root_of(nt, -1);                  // representative -> root index
    for (int i = 0; i < nt; ++i) {
        const int r = find(i);
        if (fixed[i]) {
            if (root_of[r] >= 0 && fixed[root_of[r]]) {
                P.error = "tie constraints: two prescribed DOFs are tied together (DOFs " + std::to_string(P.tdof[root_of[r]]) + " and " +
                          std::to_string(P.tdof[i]) + ", 0-based)";
                return false;
            }
            root_of[r] = i;
        } else if (root_of[r] < 0) root_of[r] = i;
    }
    std::vector<int> gid(nt, -1);
    for (int i = 0; i < nt; ++i) {
        const int r = find(i);
        if (gid[r] < 0) { gid[r] = (int)P.g_root.size(); P.g_root.push_back(P.tdof[root_of[r]]); P.g_fixed.push_back(fixed[root_of[r]]); }
    }
    std::vector<std::vector<int>> mem(P.g_root.size());
    for (int i = 0; i < nt; ++i) { const int r = find(i); if (root_of[r] != i) mem[gid[r]].push_back(P.tdof[i]); }
    for (size_t g = 0; g < P.g_root.size(); ++g) {
        for (int d : mem[g]) { P.m_dof.push_back(d); P.m_root.push_back(P.g_root[g]); P.m_fixed.push_back(P.g_fixed[g]); }
        P.g_ptr.push_back((int)P.m_dof.size());
    }
    std::vector<int> deg(nt, 0);
    std::vector<std::vector<int>> at(nt);
    for (int k = 0; k < n; ++k) { at[P.idx1[k]].push_back(k); at[P.idx2[k]].push_back(k); deg[P.idx1[k]]++; deg[P.idx2[k]]++; }
    std::vector<uint8_t> done(n, 0);
    std::vector<int> queue;
    for (int i = 0; i < nt; ++i) if (deg[i] == 1 && !fixed[i]) queue.push_back(i);
    for (size_t q = 0; q < queue.size(); ++q) {
        const int i = queue[q];
        if (deg[i] != 1) continue;
        int k = -1;
        for (int kk : at[i]) if (!done[kk]) { k = kk; break; }
        if (k < 0) continue;
        done[k] = 1;
        P.peel_tie.push_back(k); P.peel_leaf.push_back(i);
        const int o = P.idx1[k] == i ? P.idx2[k] : P.idx1[k];
        deg[i]--; deg[o]--;
        if (deg[o] == 1 && !fixed[o]) queue.push_back(o);
    }
    if ((int)P.peel_tie.size() != n) { P.error = "tie constraints: the multipliers cannot be recovered (inconsistent constraint graph)"; return false; }
    return true;
}

// the plan without a device (tests of the host logic): sizes[4] = groups, members, touched DOFs, constraints; arrays may be null
extern "C" int en234gpu_tie_plan(int n_ties, const int* dof1, const int* dof2, int n_fixed, const int* fixed_dof, int* sizes,
                                 int* group_root, int* group_ptr, int* group_fixed, int* member_dof, int* peel_tie, int* peel_dof) {
    if (n_ties <= 0 || !dof1 || !dof2 || !sizes) return set_err("en234gpu_tie_plan: bad arguments");
    TiePlan P;
    const std::vector<int> fx(fixed_dof, fixed_dof + (n_fixed > 0 ? n_fixed : 0));
    if (!tie_plan(n_ties, dof1, dof2, fx, P)) return set_err(P.error);
    sizes[0] = (int)P.g_root.size(); sizes[1] = (int)P.m_dof.size(); sizes[2] = (int)P.tdof.size(); sizes[3] = n_ties;
    for (size_t g = 0; g < P.g_root.size(); ++g) {
        if (group_root) group_root[g] = P.g_root[g];
        if (group_fixed) group_fixed[g] = P.g_fixed[g];
    }
    if (group_ptr) for (size_t g = 0; g < P.g_ptr.size(); ++g) group_ptr[g] = P.g_ptr[g];
    if (member_dof) for (size_t i = 0; i < P.m_dof.size(); ++i) member_dof[i] = P.m_dof[i];
    for (int q = 0; q < n_ties; ++q) {
        if (peel_tie) peel_tie[q] = P.peel_tie[q];
        if (peel_dof) peel_dof[q] = P.tdof[P.peel_leaf[q]];
    }
    return 0;
}

static int ties_rebuild_groups(en234gpu_ctx* c) {
    Ties& t = c->ties;
    TiePlan P;
    if (!tie_plan(t.n, t.d1.data(), t.d2.data(), c->h_fixdof, P)) return set_err(P.error);
    const int counts[2] = {(int)P.g_root.size(), (int)P.m_dof.size()};
    t.n_groups = counts[0]; t.n_members = counts[1];
    CK(cudaMemcpyAsync(t.counts.p, counts, sizeof(counts), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(t.g_root.p, P.g_root.data(), P.g_root.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(t.g_ptr.p, P.g_ptr.data(), P.g_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(t.g_fixed.p, P.g_fixed.data(), P.g_fixed.size(), cudaMemcpyHostToDevice, c->stream));
    if (!P.m_dof.empty()) {
        CK(cudaMemcpyAsync(t.m_dof.p, P.m_dof.data(), P.m_dof.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(t.m_root.p, P.m_root.data(), P.m_root.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(t.m_fixed.p, P.m_fixed.data(), P.m_fixed.size(), cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    t.peel_tie = P.peel_tie; t.peel_leaf = P.peel_leaf;
    t.groups_valid = true;
    t.groups_fixdof = c->h_fixdof;
    return 0;
}

extern "C" int en234gpu_set_ties(en234gpu_handle c, int n_ties, const int* dof1, const int* dof2) {
    CK(cudaSetDevice(c->device));
    if (n_ties < 0 || (n_ties > 0 && (!dof1 || !dof2))) return set_err("en234gpu_set_ties: bad arguments");
    if (c->comm.active) return set_err("en234gpu_set_ties: tie constraints are implemented for one device (unpartitioned handle)");
    Ties& t = c->ties;
    for (auto& g : c->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }      // the PCG / BiCGSTAB iteration changes shape
    if (c->bi_graph) { cudaGraphExecDestroy(c->bi_graph); c->bi_graph = nullptr; }
    t.reset();
    if (n_ties == 0) return 0;
    for (int k = 0; k < n_ties; ++k) {
        if (dof1[k] < 0 || dof1[k] >= c->ndof || dof2[k] < 0 || dof2[k] >= c->ndof) return set_err("en234gpu_set_ties: DOF out of range");
        if (dof1[k] == dof2[k]) return set_err("en234gpu_set_ties: a DOF is tied to itself");
    }
    t.n = n_ties;
    t.d1.assign(dof1, dof1 + n_ties); t.d2.assign(dof2, dof2 + n_ties);
    t.lambda.assign(n_ties, 0.0);
    t.h_tdof = t.d1;
    t.h_tdof.insert(t.h_tdof.end(), t.d2.begin(), t.d2.end());
    std::sort(t.h_tdof.begin(), t.h_tdof.end());
    t.h_tdof.erase(std::unique(t.h_tdof.begin(), t.h_tdof.end()), t.h_tdof.end());
    t.n_tdof = (int)t.h_tdof.size();
    auto index_of = [&](int d) { return (int)(std::lower_bound(t.h_tdof.begin(), t.h_tdof.end(), d) - t.h_tdof.begin()); };
    t.idx1.resize(n_ties); t.idx2.resize(n_ties);
    std::vector<std::vector<int>> ent(t.n_tdof);
    for (int k = 0; k < n_ties; ++k) {
        t.idx1[k] = index_of(t.d1[k]); t.idx2[k] = index_of(t.d2[k]);
        ent[t.idx1[k]].push_back(2 * k);            // row of dof1: - lambda
        ent[t.idx2[k]].push_back(2 * k + 1);        // row of dof2: + lambda
    }
    std::vector<int> tptr{0}, tent;
    for (auto& e : ent) { tent.insert(tent.end(), e.begin(), e.end()); tptr.push_back((int)tent.size()); }
    const size_t nt = (size_t)t.n_tdof, nd = (size_t)c->ndof;
    CK(t.dd1.alloc(n_ties)); CK(t.dd2.alloc(n_ties)); CK(t.dlambda.alloc(n_ties)); CK(t.gap.alloc(n_ties));
    CK(t.tdof.alloc(nt)); CK(t.tptr.alloc(nt + 1)); CK(t.tent.alloc(tent.size())); CK(t.trow.alloc(nt));
    CK(t.counts.alloc(2)); CK(t.g_root.alloc(nt)); CK(t.g_ptr.alloc(nt + 1)); CK(t.g_fixed.alloc(nt));
    CK(t.m_dof.alloc(nt)); CK(t.m_root.alloc(nt)); CK(t.m_fixed.alloc(nt)); CK(t.mgap.alloc(nt));
    CK(t.rhs0.alloc(nd)); CK(t.vec.alloc(nd)); CK(t.d0.alloc(nd));
    CK(cudaMemcpy(t.dd1.p, t.d1.data(), n_ties * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t.dd2.p, t.d2.data(), n_ties * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(t.dlambda.p, 0, n_ties * sizeof(double)));
    CK(cudaMemcpy(t.tdof.p, t.h_tdof.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t.tptr.p, tptr.data(), tptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t.tent.p, tent.data(), tent.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (ties_rebuild_groups(c)) { t.reset(); return 1; }
    return 0;
}

// a new analysis on the same handle starts from zero multipliers (read_input_file.f90:139)
extern "C" int en234gpu_reset_lagrange_multipliers(en234gpu_handle c) {
    if (c->ties.n == 0) return 0;
    CK(cudaSetDevice(c->device));
    std::fill(c->ties.lambda.begin(), c->ties.lambda.end(), 0.0);
    CK(cudaMemsetAsync(c->ties.dlambda.p, 0, c->ties.n * sizeof(double), c->stream));
    return 0;
}

extern "C" int en234gpu_get_lagrange_multipliers(en234gpu_handle c, int n, double* out) {
    if (n != c->ties.n) return set_err("en234gpu_get_lagrange_multipliers: the handle has " + std::to_string(c->ties.n) + " tie constraints");
    for (int k = 0; k < n; ++k) out[k] = c->ties.lambda[k];
    return 0;
}

static int ties_check_mode(en234gpu_ctx* c) {
    if (c->mode != EN234_SOLVE_PCG_BSR) return set_err("tie constraints: assembled operator only (EN234_SOLVE_PCG_BSR / EN234_SOLVE_BICGSTAB_BSR)");
    if (!c->diag_valid) return set_err("tie constraints need the two-kernel assembly (unset EN234_ASM=rows)");
    return 0;
}

// sum of (gap - eps lambda)^2 over the multiplier rows (their right-hand side, solver_direct.f90:325-326)
static int ties_row_norm2(en234gpu_ctx* c, bool after_bc, double* out) {
    Ties& t = c->ties;
    k_tie_gap<<<(t.n + 255) / 256, 256, 0, c->stream>>>(t.n, t.dd1.p, t.dd2.p, c->ut.p, c->du.p, after_bc ? c->cflag.p : nullptr, c->cval.p, t.gap.p);
    c->st.kernel_launches++;
    t.h_gap.resize(t.n);
    CK(cudaMemcpyAsync(t.h_gap.data(), t.gap.p, t.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double s = 0.0;
    for (int k = 0; k < t.n; ++k) { const double g = t.h_gap[k] - t.lambda[k] * t.eps; s += g * g; }
    *out = s;
    return 0;
}

// after the element assembly: multiplier terms of the tied rows, reaction forces, generalized force norm over all equations
static int ties_after_assemble(en234gpu_ctx* c, double* nodal_force_norm) {
    Ties& t = c->ties;
    if (ties_check_mode(c)) return 1;
    // eps = 1e-12 |diag| / neq over the assembled diagonal (multiplier rows are still zero there), solver_direct.f90:292
    k_sumsq<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, c->diag.p, c->part2.p);
    double dn = 0.0;
    if (finish_norm(c, c->part2.p, c->grid_vec, &dn)) return 1;
    t.eps = 1.e-12 * dn / (double)(c->ndof + t.n);
    k_tie_rhs<<<(t.n_tdof + 255) / 256, 256, 0, c->stream>>>(t.n_tdof, t.tdof.p, t.tptr.p, t.tent.p, t.dlambda.p, c->rhs.p, c->rforce.p);
    k_sumsq<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, c->rhs.p, c->part2.p);
    c->st.kernel_launches += 3;
    double rn = 0.0, g2 = 0.0;
    if (finish_norm(c, c->part2.p, c->grid_vec, &rn)) return 1;
    if (ties_row_norm2(c, false, &g2)) return 1;
    *nodal_force_norm = std::sqrt(rn * rn + g2);
    return 0;
}

static int ties_after_bc(en234gpu_ctx* c, double* unbalanced) {
    double g2 = 0.0;
    if (ties_row_norm2(c, true, &g2)) return 1;
    *unbalanced = std::sqrt(*unbalanced * *unbalanced + g2);
    return 0;
}

// before the PCG loop: keep the reference-form right-hand side, move the known member corrections to the right-hand side,
// reduce right-hand side and Jacobi diagonal to the roots
static int ties_before_solve(en234gpu_ctx* c) {
    Ties& t = c->ties;
    if (ties_check_mode(c)) return 1;
    if (!t.groups_valid || t.groups_fixdof != c->h_fixdof) {
        if (ties_rebuild_groups(c)) return 1;
    }
    const TieTables T = tie_tables(c);
    const size_t b = (size_t)c->ndof * sizeof(double);
    CK(cudaMemcpyAsync(t.rhs0.p, c->rhs.p, b, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemsetAsync(t.d0.p, 0, b, c->stream));
    const int gm = std::max(1, (t.n_tdof + 255) / 256);
    k_tie_d0<<<gm, 256, 0, c->stream>>>(T, c->ut.p, c->du.p, c->cval.p, t.d0.p, t.mgap.p);
    int np = 0;
    launch_spmv_rows(c, t.d0.p, t.vec.p, 0, c->N, 0, 0, &np, 0, nullptr, false);
    k_vec_sub<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, t.vec.p, c->rhs.p);
    k_tie_reduce_rhs<<<gm, 256, 0, c->stream>>>(T, c->rhs.p, c->minv.p);
    c->st.kernel_launches += 3;
    CK(cudaGetLastError());
    return 0;
}

// after the PCG loop: corrections of the members, multiplier corrections from the tied rows of rhs - K sol; *dl2 = their sum of squares
static int ties_after_solve(en234gpu_ctx* c, double* sol, double* dl2) {
    Ties& t = c->ties;
    const TieTables T = tie_tables(c);
    const int gm = std::max(1, (t.n_tdof + 255) / 256);
    k_tie_expand_sol<<<gm, 256, 0, c->stream>>>(T, t.mgap.p, sol);
    int np = 0;
    launch_spmv_rows(c, sol, t.vec.p, 0, c->N, 0, 0, &np, 0, nullptr, false);
    k_tie_rows<<<gm, 256, 0, c->stream>>>(t.n_tdof, t.tdof.p, t.rhs0.p, t.vec.p, t.trow.p);
    c->st.kernel_launches += 2;
    std::vector<double> row(t.n_tdof);
    CK(cudaMemcpyAsync(row.data(), t.trow.p, row.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // row of dof1: K du + dlambda = rhs;  row of dof2: K du - dlambda = rhs.  Leaves first; a resolved multiplier is moved to the
    // other end's row.
    double s = 0.0;
    for (size_t q = 0; q < t.peel_tie.size(); ++q) {
        const int k = t.peel_tie[q], leaf = t.peel_leaf[q];
        const bool first = t.idx1[k] == leaf;
        const double dl = first ? row[leaf] : -row[leaf];
        const int other = first ? t.idx2[k] : t.idx1[k];
        row[other] += first ? dl : -dl;          // remove this multiplier from the other row: (-dl) there if other is dof2, (+dl) if dof1
        t.lambda[k] += dl;
        s += dl * dl;
    }
    CK(cudaMemcpyAsync(t.dlambda.p, t.lambda.data(), t.n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *dl2 = s;
    return 0;
}
