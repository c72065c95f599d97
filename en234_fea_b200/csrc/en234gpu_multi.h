// Several GPUs behind ONE handle in ONE process (included at the end of en234gpu.cu).
// The reference's driver is a single serial program; a Fortran driver that binds this library should not have to become an
// MPI program to use eight GPUs.  en234gpu_create_multi takes the WHOLE mesh, partitions it (en234gpu_partition.h), creates
// one ordinary context per part — each driven by its own host thread, exactly as a rank of the one-process-per-GPU mode is
// — and wires the parts together through peer memory (cudaDeviceEnablePeerAccess; no IPC, no NCCL, no torch).  The
// en234gpu_multi_* entry points have the argument lists of their single-GPU namesakes with whole-mesh arrays: the library
// scatters them to the parts and gathers the results.  Inside, every part runs the same partitioned code path (halo sums,
// all-reduces folded into the PCG kernels) as a rank does.
// Several parts may name the same device (tests on a one-GPU box): the persistent grids are then sized so that the kernels
// of all parts on the device are co-resident — kernels of different parts wait for one another.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "en234gpu_partition.h"

struct en234gpu_multi_ctx {
    int P = 0, N = 0, E = 0, ndof = 0;
    std::vector<int> dev;
    std::vector<en234gpu_ctx*> ctx;
    std::vector<en234part::Part> part;
    std::vector<std::vector<int>> g2l;            // per part: global node -> local node or -1
    // host staging per part (rank-local images of the caller's arrays)
    struct Stage { std::vector<double> ut, du, fe, fx, rf, fv; std::vector<int> fd; };
    std::vector<Stage> stage;
    // worker threads: one per part
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;
    long gen = 0;
    int pending = 0;
    bool quit = false;
    std::vector<int> rc;
    std::vector<std::string> err;

    int run(std::function<int(int)> f) {
        {
            std::unique_lock<std::mutex> lk(mu);
            job = std::move(f);
            pending = P;
            ++gen;
        }
        cv_go.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        for (int p = 0; p < P; ++p)
            if (rc[p]) { g_err = "part " + std::to_string(p) + ": " + err[p]; return 1; }
        return 0;
    }
    void worker(int p) {
        long seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return quit || gen != seen; });
                if (quit) return;
                seen = gen;
                f = job;
            }
            g_err.clear();
            int r = 1;
            if (cudaSetDevice(dev[p]) == cudaSuccess) r = f(p);
            else g_err = "cudaSetDevice failed";
            {
                std::unique_lock<std::mutex> lk(mu);
                rc[p] = r;
                err[p] = g_err;
                --pending;
            }
            cv_done.notify_all();
        }
    }
};

namespace {
// rank-local image of a whole-mesh DOF array
inline void gather_dofs(const std::vector<int>& l2g, const double* global, std::vector<double>& local) {
    local.resize(3 * l2g.size());
    for (size_t i = 0; i < l2g.size(); ++i) {
        const double* s = global + 3 * (size_t)l2g[i];
        local[3 * i] = s[0]; local[3 * i + 1] = s[1]; local[3 * i + 2] = s[2];
    }
}
// owned entries only: every global entry is written by exactly one part (the copies are bitwise identical anyway)
inline void scatter_owned(const en234part::Part& P, const std::vector<double>& local, double* global) {
    for (size_t i = 0; i < P.l2g.size(); ++i)
        if (P.owned[i]) {
            double* d = global + 3 * (size_t)P.l2g[i];
            d[0] = local[3 * i]; d[1] = local[3 * i + 1]; d[2] = local[3 * i + 2];
        }
}
inline void localize_fixed(const std::vector<int>& g2l, int n_fixed, const int* dof, const double* val, std::vector<int>& ld,
                           std::vector<double>& lv) {
    ld.clear(); lv.clear();
    for (int k = 0; k < n_fixed; ++k) {
        const int n = g2l[dof[k] / 3];
        if (n >= 0) { ld.push_back(3 * n + dof[k] % 3); lv.push_back(val[k]); }
    }
}
}  // namespace

// With CUDA's lazy module loading a kernel is loaded at its first launch, and loading may need a synchronisation of the
// whole context: a part that launches some kernel for the first time would then wait for another part's kernel on the
// same device — which may be spinning for exactly this part.  Every kernel a partitioned static step can launch is
// therefore loaded before the parts start to depend on one another.
static int preload_partitioned_kernels() {
    cudaFuncAttributes a;
#define PRELOAD(...)                                                                                          \
    if (cudaFuncGetAttributes(&a, __VA_ARGS__) != cudaSuccess) return set_err("cannot load kernel " #__VA_ARGS__)
    PRELOAD(k_element_blocks<false, true, LoadSum>); PRELOAD(k_element_blocks<true, true, LoadSum>);
    PRELOAD(k_element_blocks<true, false, LoadSum>); PRELOAD(k_assemble_rows<false>); PRELOAD(k_assemble_rows<true>);
    PRELOAD(k_element_forces<LoadSum, 0>); PRELOAD(k_element_forces<LoadSum, 1>); PRELOAD(k_element_forces<LoadSum, 2>);
    PRELOAD(k_element_forces_tpe<LoadSum>); PRELOAD(k_element_forces_tpe<LoadMasked>); PRELOAD(k_element_forces_cached<LoadSum, 2>); PRELOAD(k_element_forces_cached<LoadMasked, 2>);
    PRELOAD(k_element_forces<LoadMasked, 0>); PRELOAD(k_element_forces<LoadMasked, 1>); PRELOAD(k_element_forces<LoadMasked, 2>);
    PRELOAD(k_gather_rows); PRELOAD(k_gather_residual); PRELOAD(k_negate_norm); PRELOAD(k_reduce_to); PRELOAD(k_reduce2_to);
    PRELOAD(k_pack2); PRELOAD(k_unpack2); PRELOAD(k_set_corrections); PRELOAD(k_apply_bc); PRELOAD(k_bc_rest); PRELOAD(k_bc_finalize);
    PRELOAD(k_mask_values); PRELOAD(k_mf_color<1>); PRELOAD(k_fix_diag); PRELOAD(k_mf_gather); PRELOAD(k_mask_soa);
    PRELOAD(k_stencil_nodes); PRELOAD(k_element_batch);
    PRELOAD(k_halo_push); PRELOAD(k_halo_wait_add); PRELOAD(k_halo_wait_total); PRELOAD(k_p2p_allreduce);
    PRELOAD(k_p2p_allreduce_vals); PRELOAD(k_p2p_push);
    PRELOAD(k_cg_init); PRELOAD(k_cg_init_finish); PRELOAD(k_cg_update<true>); PRELOAD(k_cg_update<false>); PRELOAD(k_cg_pupdate);
    PRELOAD(k_add_solution); PRELOAD(k_spmv<4, false, true>); PRELOAD(k_spmv<4, true, true>); PRELOAD(k_spmv<4, false>);
    PRELOAD(k_spmv<4>); PRELOAD(k_spmv<5>);
#undef PRELOAD
    return 0;
}

extern "C" {

int en234gpu_multi_destroy(en234gpu_multi_handle m) {
    if (!m) return 0;
    if (!m->th.empty()) {
        m->run([&](int p) { if (m->ctx[p]) en234gpu_destroy(m->ctx[p]); m->ctx[p] = nullptr; return 0; });
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->quit = true;
        }
        m->cv_go.notify_all();
        for (auto& t : m->th) t.join();
    }
    delete m;
    return 0;
}

int en234gpu_create_multi(en234gpu_multi_handle* out, int n_parts, const int* devices, int n_nodes, int n_elements,
                          const double* coords, const int* connectivity, const int* element_flag, const int* element_prop_index,
                          const double* element_properties, int n_element_properties) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err("en234gpu_create_multi: no CUDA device available (this library has no CPU fallback)");
    if (n_parts < 1 || n_parts > P2P_MAXR) return set_err("en234gpu_create_multi: between 1 and 16 parts");
    if (n_nodes <= 0 || n_elements <= 0) return set_err("en234gpu_create_multi: empty mesh");
    auto* m = new en234gpu_multi_ctx();
    m->P = n_parts; m->N = n_nodes; m->E = n_elements; m->ndof = 3 * n_nodes;
    for (int p = 0; p < n_parts; ++p) {
        const int d = devices ? devices[p] : p % ndev;
        if (d < 0 || d >= ndev) { delete m; return set_err("en234gpu_create_multi: bad device index"); }
        m->dev.push_back(d);
    }
    // Parts that share a device wait for one another from separately launched kernels, and nothing in CUDA guarantees that
    // those run at the same time (the library sizes the grids so that they can, preloads the kernels and preallocates, and
    // every wait has a time-out) — a debugging aid for small meshes on a one-GPU box, refused unless asked for.
    for (int p = 0; p < n_parts; ++p)
        for (int q = 0; q < p; ++q)
            if (m->dev[p] == m->dev[q]) {
                const char* e = getenv("EN234_ALLOW_SHARED_DEVICE");
                if (!(e && atoi(e) != 0)) {
                    delete m;
                    return set_err("en234gpu_create_multi: parts " + std::to_string(q) + " and " + std::to_string(p) +
                                   " are both on device " + std::to_string(m->dev[p]) +
                                   "; one part per GPU (set EN234_ALLOW_SHARED_DEVICE=1 to co-schedule parts on one device: small meshes, debugging only)");
                }
            }
    const std::string perr = en234part::partition(n_nodes, n_elements, connectivity, n_parts, m->part);
    if (!perr.empty()) { delete m; return set_err("en234gpu_create_multi: " + perr); }
    m->g2l.resize(n_parts);
    m->stage.resize(n_parts);
    m->ctx.assign(n_parts, nullptr);
    m->rc.assign(n_parts, 0);
    m->err.assign(n_parts, "");
    int max_shared = 1;
    for (auto& P : m->part)
        for (size_t k = 0; k + 1 < P.shared_ptr.size(); ++k) max_shared = std::max(max_shared, P.shared_ptr[k + 1] - P.shared_ptr[k]);
    // peer access between every pair of distinct devices
    for (int p = 0; p < n_parts; ++p)
        for (int q = 0; q < n_parts; ++q) {
            if (m->dev[p] == m->dev[q]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[p], m->dev[q]);
            if (!can) { delete m; return set_err("en234gpu_create_multi: devices " + std::to_string(m->dev[p]) + " and " + std::to_string(m->dev[q]) + " have no peer access"); }
            cudaSetDevice(m->dev[p]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { delete m; return set_err(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    for (int p = 0; p < n_parts; ++p) m->th.emplace_back(&en234gpu_multi_ctx::worker, m, p);
    int rc = m->run([&](int p) {
        const en234part::Part& P = m->part[p];
        const int nl = (int)P.l2g.size(), el = (int)(P.e1 - P.e0);
        std::vector<double> xl(3 * (size_t)nl);
        std::vector<int>& g2l = m->g2l[p];
        g2l.assign(n_nodes, -1);
        for (int i = 0; i < nl; ++i) {
            g2l[P.l2g[i]] = i;
            for (int c = 0; c < 3; ++c) xl[3 * (size_t)i + c] = coords[3 * (size_t)P.l2g[i] + c];
        }
        if (preload_partitioned_kernels()) return 1;
        int share = 0;
        for (int q = 0; q < n_parts; ++q) share += m->dev[q] == m->dev[p] ? 1 : 0;
        g_sm_share = share;
        const int r = en234gpu_create(&m->ctx[p], m->dev[p], nl, el, xl.data(), P.conn.data(), element_flag + P.e0,
                                      element_prop_index ? element_prop_index + P.e0 : nullptr, element_properties, n_element_properties);
        g_sm_share = 1;
        if (r) return r;
        en234gpu_ctx* c = m->ctx[p];
        if (n_parts == 1) return 0;
        if (en234gpu_set_partition(c, (int)P.nbr.size(), P.nbr.data(), P.shared_ptr.data(), P.shared_nodes.data(), P.owned.data())) return 1;
        if (comm_init_inproc(c->comm, n_parts, p)) return set_err(comm_error());
        if (comm_arena_create(c->comm, max_shared, nullptr)) return set_err(comm_error());
        return 0;
    });
    if (!rc && n_parts > 1) {
        std::vector<void*> bases(n_parts);
        for (int p = 0; p < n_parts; ++p) bases[p] = m->ctx[p]->comm.arena;
        rc = m->run([&](int p) {
            en234gpu_ctx* c = m->ctx[p];
            if (comm_arena_connect_inproc(c->comm, bases.data())) return set_err(comm_error());
            return 0;
        });
    }
    if (rc) { const std::string keep = g_err; en234gpu_multi_destroy(m); g_err = keep; return 1; }
    *out = m;
    return 0;
}

// the partition en234gpu_create_multi uses, for inspection and tests (host only: no device needed)
int en234gpu_partition_sizes(int n_parts, int part, int n_nodes, int n_elements, const int* connectivity, int* sizes) {
    std::vector<en234part::Part> parts;
    const std::string err = en234part::partition(n_nodes, n_elements, connectivity, n_parts, parts);
    if (!err.empty()) return set_err(err);
    if (part < 0 || part >= n_parts) return set_err("partition: no such part");
    const en234part::Part& P = parts[part];
    sizes[0] = (int)P.l2g.size(); sizes[1] = (int)(P.e1 - P.e0); sizes[2] = (int)P.e0; sizes[3] = P.n_interface;
    sizes[4] = (int)P.nbr.size(); sizes[5] = (int)P.shared_nodes.size();
    return 0;
}
int en234gpu_partition_arrays(int n_parts, int part, int n_nodes, int n_elements, const int* connectivity, int* local_to_global,
                              int* local_connectivity, int* neighbour_rank, int* shared_ptr, int* shared_nodes,
                              unsigned char* owned) {
    std::vector<en234part::Part> parts;
    const std::string err = en234part::partition(n_nodes, n_elements, connectivity, n_parts, parts);
    if (!err.empty()) return set_err(err);
    if (part < 0 || part >= n_parts) return set_err("partition: no such part");
    const en234part::Part& P = parts[part];
    std::copy(P.l2g.begin(), P.l2g.end(), local_to_global);
    std::copy(P.conn.begin(), P.conn.end(), local_connectivity);
    std::copy(P.nbr.begin(), P.nbr.end(), neighbour_rank);
    std::copy(P.shared_ptr.begin(), P.shared_ptr.end(), shared_ptr);
    std::copy(P.shared_nodes.begin(), P.shared_nodes.end(), shared_nodes);
    std::copy(P.owned.begin(), P.owned.end(), owned);
    return 0;
}

int en234gpu_multi_n_parts(en234gpu_multi_handle m) { return m->P; }
en234gpu_handle en234gpu_multi_part_handle(en234gpu_multi_handle m, int part) { return part >= 0 && part < m->P ? m->ctx[part] : nullptr; }

int en234gpu_multi_part_info(en234gpu_multi_handle m, int part, int* device, int* n_nodes, int* n_elements, int* n_interface,
                             int* n_neighbours) {
    if (part < 0 || part >= m->P) return set_err("en234gpu_multi_part_info: no such part");
    const en234part::Part& P = m->part[part];
    if (device) *device = m->dev[part];
    if (n_nodes) *n_nodes = (int)P.l2g.size();
    if (n_elements) *n_elements = (int)(P.e1 - P.e0);
    if (n_interface) *n_interface = P.n_interface;
    if (n_neighbours) *n_neighbours = (int)P.nbr.size();
    return 0;
}

int en234gpu_multi_set_operator_mode(en234gpu_multi_handle m, int method) {
    return m->run([&](int p) { return en234gpu_set_operator_mode(m->ctx[p], method); });
}

int en234gpu_multi_assemble(en234gpu_multi_handle m, const double* dof_total, const double* dof_increment,
                            const double* f_element_ext, double* rforce, double* nodal_force_norm, int* fail) {
    if (!dof_total || !dof_increment) return set_err("en234gpu_multi_assemble: null DOF arrays");
    std::vector<double> nrm(m->P, 0.0);
    std::vector<int> fl(m->P, 0);
    const int rc = m->run([&](int p) {
        const en234part::Part& P = m->part[p];
        auto& S = m->stage[p];
        gather_dofs(P.l2g, dof_total, S.ut);
        gather_dofs(P.l2g, dof_increment, S.du);
        if (f_element_ext) gather_dofs(P.l2g, f_element_ext, S.fe);
        S.rf.resize(S.ut.size());
        if (en234gpu_assemble(m->ctx[p], S.ut.data(), S.du.data(), f_element_ext ? S.fe.data() : nullptr, rforce ? S.rf.data() : nullptr,
                              &nrm[p], &fl[p])) return 1;
        if (rforce) scatter_owned(P, S.rf, rforce);
        return 0;
    });
    if (rc) return 1;
    *nodal_force_norm = nrm[0];
    *fail = 0;
    for (int f : fl) *fail |= f;
    return 0;
}

int en234gpu_multi_apply_boundaryconditions(en234gpu_multi_handle m, const double* f_ext, int n_fixed, const int* fixed_dof,
                                            const double* fixed_correction, double* unbalanced_force_norm) {
    for (int k = 0; k < n_fixed; ++k)
        if (fixed_dof[k] < 0 || fixed_dof[k] >= m->ndof) return set_err("prescribed DOF index out of range");
    std::vector<double> nrm(m->P, 0.0);
    const int rc = m->run([&](int p) {
        auto& S = m->stage[p];
        if (f_ext) gather_dofs(m->part[p].l2g, f_ext, S.fx);
        localize_fixed(m->g2l[p], n_fixed, fixed_dof, fixed_correction, S.fd, S.fv);
        return en234gpu_apply_boundaryconditions(m->ctx[p], f_ext ? S.fx.data() : nullptr, (int)S.fd.size(), S.fd.data(), S.fv.data(), &nrm[p]);
    });
    if (rc) return 1;
    *unbalanced_force_norm = nrm[0];
    return 0;
}

int en234gpu_multi_solve(en234gpu_multi_handle m, int method, double cg_tolerance, int max_cg_iterations, double* dof_increment,
                         double* correction_norm, int* cg_iterations, int* fail) {
    std::vector<double> nrm(m->P, 0.0);
    std::vector<int> its(m->P, 0), fl(m->P, 0);
    const int rc = m->run([&](int p) {
        const en234part::Part& P = m->part[p];
        auto& S = m->stage[p];
        gather_dofs(P.l2g, dof_increment, S.du);
        return en234gpu_solve(m->ctx[p], method, cg_tolerance, max_cg_iterations, S.du.data(), &nrm[p], &its[p], &fl[p]);
    });
    if (rc) return 1;
    // results are written after every part has read its input image (owned entries only: one writer per entry)
    for (int p = 0; p < m->P; ++p) scatter_owned(m->part[p], m->stage[p].du, dof_increment);
    *correction_norm = nrm[0];
    *cg_iterations = its[0];
    *fail = 0;
    for (int f : fl) *fail |= f;
    return 0;
}

// ---- device-resident variants
int en234gpu_multi_upload_state(en234gpu_multi_handle m, const double* dof_total, const double* dof_increment,
                                const double* f_element_ext, const double* f_ext, int n_fixed, const int* fixed_dof,
                                const double* fixed_value) {
    for (int k = 0; k < n_fixed; ++k)
        if (fixed_dof[k] < 0 || fixed_dof[k] >= m->ndof) return set_err("prescribed DOF index out of range");
    return m->run([&](int p) {
        const en234part::Part& P = m->part[p];
        auto& S = m->stage[p];
        if (dof_total) gather_dofs(P.l2g, dof_total, S.ut);
        if (dof_increment) gather_dofs(P.l2g, dof_increment, S.du);
        if (f_element_ext) gather_dofs(P.l2g, f_element_ext, S.fe);
        if (f_ext) gather_dofs(P.l2g, f_ext, S.fx);
        if (n_fixed >= 0) localize_fixed(m->g2l[p], n_fixed, fixed_dof, fixed_value, S.fd, S.fv);
        return en234gpu_upload_state(m->ctx[p], dof_total ? S.ut.data() : nullptr, dof_increment ? S.du.data() : nullptr,
                                     f_element_ext ? S.fe.data() : nullptr, f_ext ? S.fx.data() : nullptr,
                                     n_fixed >= 0 ? (int)S.fd.size() : -1, S.fd.data(), S.fv.data());
    });
}

int en234gpu_multi_assemble_dev(en234gpu_multi_handle m, double* nodal_force_norm, int* fail) {
    std::vector<double> nrm(m->P, 0.0);
    std::vector<int> fl(m->P, 0);
    if (m->run([&](int p) { return en234gpu_assemble_dev(m->ctx[p], &nrm[p], &fl[p]); })) return 1;
    *nodal_force_norm = nrm[0];
    *fail = 0;
    for (int f : fl) *fail |= f;
    return 0;
}

int en234gpu_multi_apply_boundaryconditions_dev(en234gpu_multi_handle m, double* unbalanced_force_norm) {
    std::vector<double> nrm(m->P, 0.0);
    if (m->run([&](int p) { return en234gpu_apply_boundaryconditions_dev(m->ctx[p], &nrm[p]); })) return 1;
    *unbalanced_force_norm = nrm[0];
    return 0;
}

int en234gpu_multi_solve_dev(en234gpu_multi_handle m, int method, double cg_tolerance, int max_cg_iterations,
                             double* correction_norm, int* cg_iterations, int* fail) {
    std::vector<double> nrm(m->P, 0.0);
    std::vector<int> its(m->P, 0), fl(m->P, 0);
    if (m->run([&](int p) { return en234gpu_solve_dev(m->ctx[p], method, cg_tolerance, max_cg_iterations, &nrm[p], &its[p], &fl[p]); })) return 1;
    *correction_norm = nrm[0];
    *cg_iterations = its[0];
    *fail = 0;
    for (int f : fl) *fail |= f;
    return 0;
}

int en234gpu_multi_zero_increment_dev(en234gpu_multi_handle m) {
    return m->run([&](int p) { return en234gpu_zero_increment_dev(m->ctx[p]); });
}

int en234gpu_multi_download_state(en234gpu_multi_handle m, double* dof_increment, double* rforce) {
    return m->run([&](int p) {
        const en234part::Part& P = m->part[p];
        auto& S = m->stage[p];
        S.du.resize(3 * P.l2g.size()); S.rf.resize(3 * P.l2g.size());
        if (en234gpu_download_state(m->ctx[p], dof_increment ? S.du.data() : nullptr, rforce ? S.rf.data() : nullptr)) return 1;
        if (dof_increment) scatter_owned(P, S.du, dof_increment);
        if (rforce) scatter_owned(P, S.rf, rforce);
        return 0;
    });
}

// device time of the last call of each phase = the slowest part's; launch counter = sum over the parts
int en234gpu_multi_get_stats(en234gpu_multi_handle m, struct en234gpu_stats* s) {
    *s = en234gpu_stats{};
    for (int p = 0; p < m->P; ++p) {
        en234gpu_stats t;
        en234gpu_get_stats(m->ctx[p], &t);
        s->assemble_ms = std::max(s->assemble_ms, t.assemble_ms);
        s->bc_ms = std::max(s->bc_ms, t.bc_ms);
        s->solve_ms = std::max(s->solve_ms, t.solve_ms);
        s->spmv_ms_avg = std::max(s->spmv_ms_avg, t.spmv_ms_avg);
        s->kernel_launches += t.kernel_launches;
        s->last_cg_iterations = t.last_cg_iterations;
        s->last_cg_err = t.last_cg_err;
        s->mf_congruent = t.mf_congruent;
    }
    return 0;
}

}  // extern "C"
