// liben234gpu.so — C ABI (include/en234gpu.h) over the sm_100a kernels.  Host code here only builds index
// structures once per analysis and launches kernels; there is no CPU implementation of any numerical step.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/en234gpu.h"
#include "en234gpu_asm.cuh"
#include "en234gpu_aux.cuh"
#include "en234gpu_bicg.cuh"
#include "en234gpu_c3d8.cuh"
#include "en234gpu_elem.cuh"
#include "en234gpu_cg.cuh"
#include "en234gpu_mf.cuh"
#include "en234gpu_project.cuh"
#include "en234gpu_dyn.cuh"
#include "en234gpu_generic.cuh"
#include "en234gpu_comm.h"
#include "en234gpu_ties.cuh"

using namespace en234k;

static thread_local std::string g_err;
static thread_local int g_sm_share = 1;   // parts that share the creating thread's device (en234gpu_create_multi)
static int set_err(const std::string& s) { g_err = s; return 1; }
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            return set_err(std::string(#call) + ": " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                           std::to_string(__LINE__));                                                \
        }                                                                                            \
    } while (0)

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

constexpr int CG_CHUNK = 32;   // PCG iterations per CUDA-graph launch (even: iteration parity is baked into the graph)

// what the host reads back after every chunk of captured iterations (pinned; two slots: one chunk is always in flight)
struct CgPoll { en234k::CgScalars S; int p2p_error; int pad; };

// two-node tie constraints (en234gpu_ties.cuh, en234gpu_ties_host.h)
struct Ties {
    int n = 0;                               // constraints
    std::vector<int> d1, d2;                 // u[d1] = u[d2], 0-based DOFs
    std::vector<double> lambda, h_gap;       // Lagrange multipliers (host copy is the master), gaps of the last norm
    std::vector<int> h_tdof, idx1, idx2;     // touched DOFs (sorted), index of d1 / d2 in that list
    int n_tdof = 0, n_groups = 0, n_members = 0;
    double eps = 0.0;                        // compliance of the multiplier rows: 1e-12 |diag| / neq
    DevBuf<int> dd1, dd2, tdof, tptr, tent, counts, g_root, g_ptr, m_dof, m_root;
    DevBuf<uint8_t> g_fixed, m_fixed;
    DevBuf<double> dlambda, gap, trow, mgap, rhs0, vec, d0;
    std::vector<int> peel_tie, peel_leaf;    // recovery order of the multiplier corrections
    bool groups_valid = false;
    std::vector<int> groups_fixdof;          // prescribed set the groups were built for
    void reset() {
        for (DevBuf<int>* b : {&dd1, &dd2, &tdof, &tptr, &tent, &counts, &g_root, &g_ptr, &m_dof, &m_root}) b->release();
        g_fixed.release(); m_fixed.release();
        for (DevBuf<double>* b : {&dlambda, &gap, &trow, &mgap, &rhs0, &vec, &d0}) b->release();
        n = n_tdof = n_groups = n_members = 0;
        d1.clear(); d2.clear(); lambda.clear(); h_gap.clear(); h_tdof.clear(); idx1.clear(); idx2.clear();
        peel_tie.clear(); peel_leaf.clear(); groups_fixdof.clear();
        groups_valid = false; eps = 0.0;
    }
};

struct en234gpu_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int N = 0, E = 0, ndof = 0, n_sm = 148;
    int npe = 8;                 // nodes per element: 8 = the hex8 path; 4 / 10 / 20 = general path (en234gpu_generic.cuh)
    DevBuf<int> gconn, gadj;     // general path: conn[e * npe + a], adjacency entries e * 32 + a
    DevBuf<uint8_t> gslot;
    long long nnzb = 0;
    // mesh
    DevBuf<double> x, y, z;
    DevBuf<int> conn, mat, adj_ptr, adj, adj_pos, browptr, bcol;
    DevBuf<uint8_t> slot;
    DevBuf<unsigned> inv;        // inverse slot map per block (k_gather_rows)
    DevBuf<Material> mats;
    // system
    DevBuf<double> val, rhs, rforce, minv, ut, du, felem, fext, cval, sol, r, p, Ap, part, part2, part3, tmpvec, tmpvec2;
    DevBuf<uint8_t> cflag, owned;
    DevBuf<int> fixdof;
    DevBuf<double> fixval;
    int n_fixed = 0;
    bool have_felem = false, have_fext = false, have_owned = false;
    DevBuf<CgScalars> S;
    CgScalars* hS = nullptr;     // pinned mirror
    CgPoll* hPoll = nullptr;     // pinned, 2 slots
    cudaEvent_t evPoll[2] = {nullptr, nullptr};
    // matrix-free
    MfData mf;
    // graphs
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    long long graph_launches[2] = {0, 0};    // kernels inside one graph launch
    // launch geometry
    int grid_asm = 0, grid_bc = 0, grid_spmv = 0, grid_vec = 0, grid_elem = 0, grid_elem0 = 0, grid_gather = 0;
    size_t asm_smem = 0, elem_smem = 0;
    int asm_threads = THREADS, elem_threads = THREADS;
    int asm_rows = 0;            // EN234_ASM=rows: single-kernel row-gather assembly (A/B and low-memory path)
    DevBuf<double> Ke, Re;       // element scratch of the two-kernel assembly (allocated at the first assembly)
    DevBuf<double> diag;         // assembled diagonal (written by k_gather_rows)
    DevBuf<int> bc_rows;         // block rows touched by the prescribed DOFs (prescribed nodes and their neighbours)
    DevBuf<uint8_t> bc_affected; // the same set as a per-node flag
    int n_bc_rows = 0;
    bool bc_rows_valid = false;  // bc_rows / bc_affected describe h_fixdof
    std::vector<int> h_fixdof;   // last prescribed-DOF list (the row subset is rebuilt only when it changes)
    bool diag_valid = false;
    int spmv_occ = 4;
    int spmv_sym = 0;            // EN234_SPMV_SYM=1: 0 untried, 1 on (index built), -1 not applicable to this matrix
    DevBuf<unsigned> sym_base;   // per block: offset of the transposed block in val (0xffffffff: read in place)
    DevBuf<uint8_t> sym_len;     // per block: scalar-row length of the row that stores the transposed block
    int max_row_blocks = 0;      // longest block row (k_spmv drops its tail loop when <= 32)
    bool any_neo = false;
    bool any_c3d8 = false;       // built-in C3D8 continuum elements (all elements of the mesh)
    DevBuf<double> sv0, sv1;     // initial / updated element state variables
    DevBuf<double> bi_z, bi_t, bi_xt, bi_b;   // BiCGSTAB work vectors
    DevBuf<BiScalars> bi_S;                   // device-resident BiCGSTAB recurrence scalars
    BiScalars* h_bi_S = nullptr;              // pinned mirror
    cudaGraphExec_t bi_graph = nullptr;       // BI_CHUNK iterations
    long long bi_graph_launches = 0;
    // explicit dynamics (en234gpu_dyn.cuh)
    struct Dyn {
        bool ready = false;
        DevBuf<double> v, a, mass, f, density, hist_t, hist_v, load_val, pd_value, time;
        DevBuf<int> hist_ptr, load_dof, pd_dof, pd_hist, pd_mode;
        std::vector<int> load_ptr, load_hist;
        std::vector<double> load_const;
        int n_pd = 0;
        cudaGraphExec_t graph = nullptr;       // DYN_CHUNK steps at time step graph_dt
        double graph_dt = 0.0;
        long long graph_launches = 0;
    } dyn;
    // host copies needed for inspection
    std::vector<int> h_browptr, h_bcol;
    // multi-GPU
    Comm comm;
    Ties ties;
    // stats
    en234gpu_stats st{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool matrix_valid = false, rhs_valid = false;
    int mode = EN234_SOLVE_PCG_BSR;   // operator mode: assembled block-CSR or matrix-free
};

static Mesh mesh_of(en234gpu_ctx* c) {
    Mesh M;
    M.N = c->N; M.E = c->E;
    M.x = c->x.p; M.y = c->y.p; M.z = c->z.p;
    M.conn = c->conn.p; M.mat = c->mat.p; M.mats = c->mats.p; M.n_mats = (int)c->mats.n;
    M.adj_ptr = c->adj_ptr.p; M.adj = c->adj.p; M.adj_pos = c->adj_pos.p; M.slot = c->slot.p;
    M.browptr = c->browptr.p; M.bcol = c->bcol.p;
    return M;
}

// Kernel launch with the programmatic-dependent-launch attribute (the kernel calls pdl_wait() before it touches what its
// predecessor in the stream wrote); opt-in with EN234_PDL=1 (64^3: 101.4 us per PCG iteration with and without it).
// Programmatic dependent launch of the PCG kernels.  One GPU: no gain inside CUDA graphs (101.4 us at 64^3 with and without).
// Partitioned loop over peer memory: its chain of short kernels gains 4.5 us per iteration on 2 B200 (116.6 -> 112.1), so it
// is on there by default.  EN234_PDL=0 / 1 overrides both.
static thread_local bool t_pdl_auto = false;      // set per captured iteration (one host thread per handle)
static bool use_pdl() {
    static const int env = getenv("EN234_PDL") ? atoi(getenv("EN234_PDL")) : -1;
    return env >= 0 ? env == 1 : t_pdl_auto;
}
// The interior-row SpMV of a partitioned run is launched programmatically only on request (EN234_PDL=1): launched early, its
// persistent CTAs would take the SM slots before the interface-row SpMV of the side stream — which everything downstream
// (halo push, the peers' waits) hangs on — has started; the short kernels of the chain have no such side effect.
static bool use_pdl_interior_spmv() {
    static const bool on = getenv("EN234_PDL") && atoi(getenv("EN234_PDL")) == 1;
    return on;
}
template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

extern "C" {

static int preallocate_for_shared_device(en234gpu_ctx* c);
const char* en234gpu_last_error(void) { return g_err.c_str(); }

int en234gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int en234gpu_create(en234gpu_handle* out, int device, int n_nodes, int n_elements, const double* coords,
                    const int* connectivity, const int* element_flag, const int* element_prop_index,
                    const double* element_properties, int n_element_properties) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err("en234gpu_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err("en234gpu_create: bad device index");
    if (n_nodes <= 0 || n_elements <= 0) return set_err("en234gpu_create: empty mesh");
    if ((long long)n_nodes * 3 >= 2147483647LL) return set_err("en234gpu_create: more than 2^31 DOFs per GPU");
    CK(cudaSetDevice(device));
    auto* c = new en234gpu_ctx();
    c->device = device;
    c->N = n_nodes; c->E = n_elements; c->ndof = 3 * n_nodes;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    // several parts of one mesh on this device (tests on a one-GPU box): their kernels wait for one another and must be
    // co-resident, so each part sizes its persistent grids for its share of half the SMs
    if (g_sm_share > 1) c->n_sm = std::max(1, c->n_sm / (2 * g_sm_share));
    const int N = n_nodes, E = n_elements;

    // ---- materials: one table entry per distinct (flag, property offset)
    std::map<std::pair<int, int>, int> matmap;
    std::vector<Material> mats;
    std::vector<int> mat(E);
    for (int e = 0; e < E; ++e) {
        const int flag = element_flag[e], pi = element_prop_index ? element_prop_index[e] : 0;
        auto key = std::make_pair(flag, pi);
        auto it = matmap.find(key);
        if (it == matmap.end()) {
            Material m{};
            if (flag == EN234_ID_LINELAST || flag == EN234_FLAG_UEL_BASE + 1) {
                if (pi + 2 > n_element_properties) { delete c; return set_err("element needs 2 properties (E, nu)"); }
                const double Em = element_properties[pi], xnu = element_properties[pi + 1];
                m.kind = 0;                                           // el_linelast_3Dbasic.f90:93-98
                m.d44 = 0.5 * Em / (1 + xnu);
                m.d11 = (1.0 - xnu) * Em / ((1 + xnu) * (1 - 2.0 * xnu));
                m.d12 = xnu * Em / ((1 + xnu) * (1 - 2.0 * xnu));
            } else if (flag == EN234_ID_NEOHOOKE || flag == EN234_FLAG_UEL_BASE + 2) {
                if (pi + 2 > n_element_properties) { delete c; return set_err("element needs 2 properties (mu, K)"); }
                m.kind = 1;
                m.mu = element_properties[pi];
                m.kappa = element_properties[pi + 1];
            } else if (flag == EN234_FLAG_C3D8) {
                if (pi + 2 > n_element_properties) { delete c; return set_err("C3D8 element needs 2 material properties (E, nu)"); }
                m.kind = 2;                                           // abaqus_umat_elastic.for:253-308
                m.d11 = element_properties[pi];                       // E
                m.d12 = element_properties[pi + 1];                   // nu
            } else {
                delete c;                                             // user_element.f90:83-86
                return set_err("invalid element type " + std::to_string(flag) +
                               " (device element types: 1001, 1002, U1, U2, C3D8)");
            }
            it = matmap.emplace(key, (int)mats.size()).first;
            mats.push_back(m);
        }
        mat[e] = it->second;
    }
    bool any_neo = false, any_c3d8 = false, any_user = false;
    for (auto& m : mats) { any_neo |= m.kind == 1; any_c3d8 |= m.kind == 2; any_user |= m.kind != 2; }
    if (any_c3d8 && any_user) { delete c; return set_err("a mesh mixing C3D8 and user elements is not supported on the device"); }

    // ---- SoA mesh, adjacency (ascending element order), block-CSR pattern, slot map
    std::vector<double> xs(N), ys(N), zs(N);
    for (int n = 0; n < N; ++n) { xs[n] = coords[3 * (size_t)n]; ys[n] = coords[3 * (size_t)n + 1]; zs[n] = coords[3 * (size_t)n + 2]; }
    std::vector<int> conn((size_t)8 * E);
    std::vector<int> adj_ptr(N + 1, 0);
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < 8; ++a) {
            const int n = connectivity[8 * (size_t)e + a] - 1;
            if (n < 0 || n >= N) { delete c; return set_err("connectivity refers to a node that does not exist"); }
            conn[(size_t)a * E + e] = n;
            adj_ptr[n + 1]++;
        }
    for (int n = 0; n < N; ++n) adj_ptr[n + 1] += adj_ptr[n];
    std::vector<int> adj((size_t)8 * E), fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < 8; ++a) adj[fill[conn[(size_t)a * E + e]]++] = e * 8 + a;
    std::vector<int>& browptr = c->h_browptr;
    std::vector<int>& bcol = c->h_bcol;
    browptr.assign(N + 1, 0);
    bcol.clear();
    bcol.reserve((size_t)27 * N);
    std::vector<uint8_t> slot((size_t)64 * E);
    std::vector<int> cols;
    for (int i = 0; i < N; ++i) {
        cols.clear();
        for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; ++k) {
            const int e = adj[k] >> 3;
            for (int b = 0; b < 8; ++b) cols.push_back(conn[(size_t)b * E + e]);
        }
        if (cols.empty()) cols.push_back(i);          // isolated node: keep a diagonal block so 1/diag is defined
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        if (cols.size() > 255) { delete c; return set_err("a node couples to more than 255 nodes"); }
        for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; ++k) {
            const int e = adj[k] >> 3;
            for (int b = 0; b < 8; ++b) {
                const int col = conn[(size_t)b * E + e];
                slot[(size_t)k * 8 + b] = (uint8_t)(std::lower_bound(cols.begin(), cols.end(), col) - cols.begin());
            }
        }
        bcol.insert(bcol.end(), cols.begin(), cols.end());
        if (bcol.size() > 2000000000ULL) { delete c; return set_err("too many matrix blocks for one GPU"); }
        browptr[i + 1] = (int)bcol.size();
    }
    c->nnzb = (long long)bcol.size();
    for (int i = 0; i < N; ++i) c->max_row_blocks = std::max(c->max_row_blocks, browptr[i + 1] - browptr[i]);
    // inverse slot map: for block (row i, slot s) nibble k = local node b of the row's k-th adjacent element on column s
    std::vector<unsigned> inv(bcol.size(), 0xffffffffu);
    for (int i = 0; i < N; ++i) {
        const int cnt = std::min(8, adj_ptr[i + 1] - adj_ptr[i]);
        for (int k = 0; k < cnt; ++k)
            for (int b = 0; b < 8; ++b) {
                unsigned& w = inv[(size_t)browptr[i] + slot[((size_t)adj_ptr[i] + k) * 8 + b]];
                w = (w & ~(15u << (4 * k))) | ((unsigned)b << (4 * k));
            }
    }

    // ---- device storage
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
#define UP(buf, vec)                                                                                  \
    CK(c->buf.alloc((vec).size()));                                                                    \
    CK(cudaMemcpy(c->buf.p, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice))
    UP(x, xs); UP(y, ys); UP(z, zs); UP(conn, conn); UP(mat, mat); UP(mats, mats);
    std::vector<int> adj_pos((size_t)8 * E);
    for (size_t k = 0; k < adj.size(); ++k) adj_pos[adj[k]] = (int)k;
    UP(adj_pos, adj_pos);
    UP(adj_ptr, adj_ptr); UP(adj, adj); UP(slot, slot); UP(browptr, browptr); UP(bcol, bcol); UP(inv, inv);
#undef UP
    const size_t nd = (size_t)c->ndof;
    CK(c->val.alloc(9 * (size_t)c->nnzb));
    for (DevBuf<double>* b : {&c->rhs, &c->rforce, &c->minv, &c->ut, &c->du, &c->felem, &c->fext, &c->cval, &c->sol, &c->r,
                              &c->p, &c->Ap, &c->tmpvec, &c->tmpvec2}) {
        CK(b->alloc(nd));
        CK(cudaMemset(b->p, 0, nd * sizeof(double)));
    }
    CK(c->cflag.alloc(nd));
    CK(cudaMemset(c->cflag.p, 0, nd));
    CK(c->part.alloc(MAX_PARTIALS)); CK(c->part2.alloc(MAX_PARTIALS)); CK(c->part3.alloc(MAX_PARTIALS));
    CK(c->S.alloc(1));
    CK(cudaMemset(c->S.p, 0, sizeof(CgScalars)));
    CK(cudaMallocHost((void**)&c->hS, sizeof(CgScalars)));
    std::memset(c->hS, 0, sizeof(CgScalars));
    CK(cudaMallocHost((void**)&c->hPoll, 2 * sizeof(CgPoll)));
    std::memset(c->hPoll, 0, 2 * sizeof(CgPoll));
    CK(cudaEventCreateWithFlags(&c->evPoll[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evPoll[1], cudaEventDisableTiming));

    // ---- launch geometry: persistent grids sized from the SM count (148 on B200)
    c->any_neo = any_neo;
    c->any_c3d8 = any_c3d8;
    if (any_c3d8) {
        CK(c->sv0.alloc((size_t)E * C3D8_NSV)); CK(c->sv1.alloc((size_t)E * C3D8_NSV));
        CK(cudaMemset(c->sv0.p, 0, (size_t)E * C3D8_NSV * sizeof(double)));
        CK(cudaMemset(c->sv1.p, 0, (size_t)E * C3D8_NSV * sizeof(double)));
    }
    c->asm_smem = any_neo ? AsmCfg<true>::SMEM : AsmCfg<false>::SMEM;
    c->asm_threads = (any_neo ? AsmCfg<true>::NW : AsmCfg<false>::NW) * 32;
    if (any_neo) CK(cudaFuncSetAttribute(k_assemble_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->asm_smem));
    else CK(cudaFuncSetAttribute(k_assemble_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->asm_smem));
    // persistent single-wave grids: SM count x resident CTAs per SM (a partial last wave leaves most of the chip idle)
    {
        int oa = 1, ob = 1, os = 1;
        if (any_neo) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oa, k_assemble_rows<true>, c->asm_threads, c->asm_smem));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oa, k_assemble_rows<false>, c->asm_threads, c->asm_smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ob, k_apply_bc, THREADS, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&os, k_spmv<4>, THREADS, 0));
        c->spmv_occ = 4;
        if (const char* e = getenv("EN234_SPMV_OCC")) {
            if (atoi(e) == 5) { c->spmv_occ = 5; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&os, k_spmv<5>, THREADS, 0)); }
        }
        const int need = (N + WARPS - 1) / WARPS;
        const int need_asm = (N + c->asm_threads / 32 - 1) / (c->asm_threads / 32);
        c->grid_asm = std::min(MAX_PARTIALS, std::max(1, std::min(need_asm, c->n_sm * std::max(oa, 1))));
        c->grid_bc = std::min(MAX_PARTIALS, std::max(1, std::min(need, c->n_sm * std::max(ob, 1))));
        c->grid_spmv = std::min(MAX_PARTIALS, std::max(1, std::min(need, c->n_sm * std::max(os, 1))));
    }
    c->grid_vec = std::min(std::min(MAX_PARTIALS, P2P_MAXCTA), std::max(1, std::min((int)((nd + VEC_THREADS - 1) / VEC_THREADS), c->n_sm * 4)));
    {
        // two-kernel assembly: element kernel (one warp per four elements) + row gather
        int oe = 1, oe0 = 1, og = 1;
        c->elem_smem = any_neo ? ElemCfg<true>::SMEM : ElemCfg<false>::SMEM;
        c->elem_threads = (any_neo ? ElemCfg<true>::NW : ElemCfg<false>::NW) * 32;
        if (any_neo) {
            CK(cudaFuncSetAttribute(k_element_blocks<true, true, LoadSum>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->elem_smem));
            CK(cudaFuncSetAttribute(k_element_blocks<true, false, LoadSum>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->elem_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oe, k_element_blocks<true, true, LoadSum>, c->elem_threads, c->elem_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oe0, k_element_blocks<true, false, LoadSum>, c->elem_threads, c->elem_smem));
        } else {
            CK(cudaFuncSetAttribute(k_element_blocks<false, true, LoadSum>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->elem_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oe, k_element_blocks<false, true, LoadSum>, c->elem_threads, c->elem_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oe0, k_element_forces<LoadSum>, 256, 0));
        }
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&og, k_gather_rows, THREADS, 0));
        const int wpc = c->elem_threads / 32, groups = (E + 3) / 4;
        c->grid_elem = std::max(1, std::min((groups + wpc - 1) / wpc, c->n_sm * std::max(oe, 1)));
        c->grid_elem0 = std::max(1, std::min((groups + wpc - 1) / wpc, c->n_sm * std::max(oe0, 1)));
        c->grid_gather = std::min(MAX_PARTIALS, std::max(1, std::min((N + WARPS - 1) / WARPS, c->n_sm * std::max(og, 1))));
        if (const char* e = getenv("EN234_ASM")) c->asm_rows = std::strcmp(e, "rows") == 0;
    }
    c->mf.sm_share = g_sm_share > 1 ? 2 * g_sm_share : 1;
    int rc = mf_setup(c->mf, mesh_of(c), conn, N, E, c->n_sm, c->ndof);
    if (rc) { delete c; return set_err("matrix-free set-up failed"); }
    // congruent mesh (every element a translate of element 0, one linear-elastic material): the matrix-free operator
    // multiplies by the single K_e of element 0, evaluated once by the element kernel.  EN234_MF_GENERAL=1 disables it.
    bool cong = !any_neo && !any_c3d8 && mats.size() == 1 && E > 1 && getenv("EN234_MF_GENERAL") == nullptr;
    for (int e = 1; e < E && cong; ++e)
        for (int a = 1; a < 8 && cong; ++a) {
            const int n0 = conn[e], na = conn[(size_t)a * E + e], m0 = conn[0], ma = conn[(size_t)a * E];
            cong = xs[na] - xs[n0] == xs[ma] - xs[m0] && ys[na] - ys[n0] == ys[ma] - ys[m0] && zs[na] - zs[n0] == zs[ma] - zs[m0];
        }
    if (cong) {
        DevBuf<int> d_el;
        DevBuf<double> d_K, d_R;
        const int zero = 0;
        std::vector<double> K0(576);
        CK(d_el.alloc(1)); CK(d_K.alloc(576)); CK(d_R.alloc(24));
        CK(cudaMemcpy(d_el.p, &zero, sizeof(int), cudaMemcpyHostToDevice));
        k_element_batch<<<1, 64, 0, c->stream>>>(mesh_of(c), 1, d_el.p, c->ut.p, c->du.p, d_K.p, d_R.p, c->S.p);
        CK(cudaMemcpyAsync(K0.data(), d_K.p, 576 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemset(c->S.p, 0, sizeof(CgScalars)));
        const int rcs = mf_set_congruent(c->mf, K0.data(), N, adj_ptr, adj, slot, browptr, bcol);
        if (rcs == 1) { delete c; return set_err("matrix-free set-up failed (congruent mesh row table)"); }
    }
    if (g_sm_share > 1 && preallocate_for_shared_device(c)) { delete c; return 1; }
    *out = c;
    return 0;
}

static GenMesh gen_mesh_of(en234gpu_ctx* c) {
    GenMesh M;
    M.N = c->N; M.E = c->E; M.npe = c->npe;
    M.x = c->x.p; M.y = c->y.p; M.z = c->z.p;
    M.conn = c->gconn.p; M.mat = c->mat.p; M.mats = c->mats.p;
    M.adj_ptr = c->adj_ptr.p; M.adj = c->gadj.p; M.slot = c->gslot.p;
    M.browptr = c->browptr.p; M.bcol = c->bcol.p;
    return M;
}

// Meshes of 4-, 10- or 20-node elements of type 1001 (8 nodes: en234gpu_create and the hex8 kernels).
int en234gpu_create_nodes(en234gpu_handle* out, int device, int nodes_per_element, int n_nodes, int n_elements,
                          const double* coords, const int* connectivity, const int* element_flag,
                          const int* element_prop_index, const double* element_properties, int n_element_properties) {
    const int npe = nodes_per_element;
    if (npe == 8)
        return en234gpu_create(out, device, n_nodes, n_elements, coords, connectivity, element_flag, element_prop_index,
                               element_properties, n_element_properties);
    *out = nullptr;
    if (npe != 4 && npe != 10 && npe != 20) return set_err("en234gpu_create_nodes: 3-D elements have 4, 10, 8 or 20 nodes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err("en234gpu_create_nodes: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err("en234gpu_create_nodes: bad device index");
    if (n_nodes <= 0 || n_elements <= 0) return set_err("en234gpu_create_nodes: empty mesh");
    if ((long long)n_nodes * 3 >= 2147483647LL || (long long)n_elements * 32 >= 2147483647LL)
        return set_err("en234gpu_create_nodes: mesh too large for 32-bit indices");
    CK(cudaSetDevice(device));
    auto* c = new en234gpu_ctx();
    c->device = device;
    c->npe = npe;
    c->N = n_nodes; c->E = n_elements; c->ndof = 3 * n_nodes;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    const int N = n_nodes, E = n_elements;
    std::map<int, int> matmap;
    std::vector<Material> mats;
    std::vector<int> mat(E);
    for (int e = 0; e < E; ++e) {
        if (element_flag[e] != EN234_ID_LINELAST) {
            delete c;
            return set_err("4-, 10- and 20-node elements are implemented for element type 1001 only (element " + std::to_string(e + 1) + ")");
        }
        const int pi = element_prop_index ? element_prop_index[e] : 0;
        auto it = matmap.find(pi);
        if (it == matmap.end()) {
            if (pi + 2 > n_element_properties) { delete c; return set_err("element needs 2 properties (E, nu)"); }
            Material m{};
            const double Em = element_properties[pi], xnu = element_properties[pi + 1];
            m.kind = 0;                                               // el_linelast_3Dbasic.f90:93-98
            m.d44 = 0.5 * Em / (1 + xnu);
            m.d11 = (1.0 - xnu) * Em / ((1 + xnu) * (1 - 2.0 * xnu));
            m.d12 = xnu * Em / ((1 + xnu) * (1 - 2.0 * xnu));
            it = matmap.emplace(pi, (int)mats.size()).first;
            mats.push_back(m);
        }
        mat[e] = it->second;
    }
    std::vector<double> xs(N), ys(N), zs(N);
    for (int n = 0; n < N; ++n) { xs[n] = coords[3 * (size_t)n]; ys[n] = coords[3 * (size_t)n + 1]; zs[n] = coords[3 * (size_t)n + 2]; }
    std::vector<int> conn((size_t)npe * E), adj_ptr(N + 1, 0);
    for (size_t q = 0; q < conn.size(); ++q) {
        const int n = connectivity[q] - 1;
        if (n < 0 || n >= N) { delete c; return set_err("connectivity refers to a node that does not exist"); }
        conn[q] = n;
        adj_ptr[n + 1]++;
    }
    for (int n = 0; n < N; ++n) adj_ptr[n + 1] += adj_ptr[n];
    std::vector<int> adj(conn.size()), fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < npe; ++a) adj[fill[conn[(size_t)e * npe + a]]++] = e * 32 + a;
    std::vector<int>& browptr = c->h_browptr;
    std::vector<int>& bcol = c->h_bcol;
    browptr.assign(N + 1, 0);
    bcol.clear();
    std::vector<uint8_t> slot((size_t)npe * adj.size());
    std::vector<int> cols;
    for (int i = 0; i < N; ++i) {
        cols.clear();
        for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; ++k) {
            const int e = adj[k] >> 5;
            for (int b = 0; b < npe; ++b) cols.push_back(conn[(size_t)e * npe + b]);
        }
        if (cols.empty()) cols.push_back(i);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        if (cols.size() > 255) { delete c; return set_err("a node couples to more than 255 nodes"); }
        for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; ++k) {
            const int e = adj[k] >> 5;
            for (int b = 0; b < npe; ++b)
                slot[(size_t)k * npe + b] = (uint8_t)(std::lower_bound(cols.begin(), cols.end(), conn[(size_t)e * npe + b]) - cols.begin());
        }
        bcol.insert(bcol.end(), cols.begin(), cols.end());
        if (bcol.size() > 2000000000ULL) { delete c; return set_err("too many matrix blocks for one GPU"); }
        browptr[i + 1] = (int)bcol.size();
    }
    c->nnzb = (long long)bcol.size();
    for (int i = 0; i < N; ++i) c->max_row_blocks = std::max(c->max_row_blocks, browptr[i + 1] - browptr[i]);
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
#define UP(buf, vec)                                                                                  \
    CK(c->buf.alloc((vec).size()));                                                                    \
    CK(cudaMemcpy(c->buf.p, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice))
    UP(x, xs); UP(y, ys); UP(z, zs); UP(gconn, conn); UP(mat, mat); UP(mats, mats);
    UP(adj_ptr, adj_ptr); UP(gadj, adj); UP(gslot, slot); UP(browptr, browptr); UP(bcol, bcol);
#undef UP
    const size_t nd = (size_t)c->ndof;
    CK(c->val.alloc(9 * (size_t)c->nnzb));
    for (DevBuf<double>* b : {&c->rhs, &c->rforce, &c->minv, &c->ut, &c->du, &c->felem, &c->fext, &c->cval, &c->sol, &c->r,
                              &c->p, &c->Ap, &c->tmpvec, &c->tmpvec2, &c->diag}) {
        CK(b->alloc(nd));
        CK(cudaMemset(b->p, 0, nd * sizeof(double)));
    }
    CK(c->cflag.alloc(nd));
    CK(cudaMemset(c->cflag.p, 0, nd));
    CK(c->part.alloc(MAX_PARTIALS)); CK(c->part2.alloc(MAX_PARTIALS)); CK(c->part3.alloc(MAX_PARTIALS));
    CK(c->S.alloc(1));
    CK(cudaMemset(c->S.p, 0, sizeof(CgScalars)));
    CK(cudaMallocHost((void**)&c->hS, sizeof(CgScalars)));
    std::memset(c->hS, 0, sizeof(CgScalars));
    CK(cudaMallocHost((void**)&c->hPoll, 2 * sizeof(CgPoll)));
    std::memset(c->hPoll, 0, 2 * sizeof(CgPoll));
    CK(cudaEventCreateWithFlags(&c->evPoll[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evPoll[1], cudaEventDisableTiming));
    CK(c->Ke.alloc((size_t)E * 9 * npe * npe));
    CK(c->Re.alloc((size_t)E * 3 * npe));
    {
        int ob = 1, os = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ob, k_apply_bc, THREADS, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&os, k_spmv<4>, THREADS, 0));
        const int need = (N + WARPS - 1) / WARPS;
        c->grid_bc = std::min(MAX_PARTIALS, std::max(1, std::min(need, c->n_sm * std::max(ob, 1))));
        c->grid_spmv = std::min(MAX_PARTIALS, std::max(1, std::min(need, c->n_sm * std::max(os, 1))));
        c->grid_gather = c->grid_bc;
    }
    c->grid_vec = std::min(std::min(MAX_PARTIALS, P2P_MAXCTA), std::max(1, std::min((int)((nd + VEC_THREADS - 1) / VEC_THREADS), c->n_sm * 4)));
    *out = c;
    return 0;
}

int en234gpu_destroy(en234gpu_handle c) {
    if (c && c->dyn.graph) { cudaGraphExecDestroy(c->dyn.graph); c->dyn.graph = nullptr; }
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& g : c->graph) if (g) cudaGraphExecDestroy(g);
    if (c->bi_graph) cudaGraphExecDestroy(c->bi_graph);
    if (c->h_bi_S) cudaFreeHost(c->h_bi_S);
    comm_destroy(c->comm);
    mf_release(c->mf);
    if (c->hS) cudaFreeHost(c->hS);
    if (c->hPoll) cudaFreeHost(c->hPoll);
    for (auto& e : c->evPoll) if (e) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

int en234gpu_set_stream(en234gpu_handle c, void* s) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    for (auto& g : c->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (c->bi_graph) { cudaGraphExecDestroy(c->bi_graph); c->bi_graph = nullptr; }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
static int read_scalars(en234gpu_ctx* c) {
    CK(cudaMemcpyAsync(c->hS, c->S.p, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
    // a peer-memory collective that timed out anywhere before this point (halo sums and all-reduces of the assembly /
    // boundary-condition phases included) is reported here, not at some later solve
    if (c->comm.p2p)
        CK(cudaMemcpyAsync(&c->hPoll[0].p2p_error, &((ArenaHdr*)c->comm.arena)->error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->comm.p2p && c->hPoll[0].p2p_error)
        return set_err("peer-memory collective timed out (a rank stopped responding); the communicator must be rebuilt");
    return 0;
}

// Several parts of one mesh on ONE device (en234gpu_create_multi with a repeated device): a device memory allocation is an
// implicit synchronisation point of the whole device, and a part whose kernel is waiting for another part's push must
// never sit behind one — so everything the steady-state calls would allocate lazily is allocated at creation.
static int preallocate_for_shared_device(en234gpu_ctx* c) {
    const size_t need_ke = (size_t)((c->E + 3) / 4) * 4 * 576;
    if (!c->asm_rows) { CK(c->Ke.alloc(need_ke)); }
    CK(c->Re.alloc((size_t)c->E * 24));
    CK(c->diag.alloc((size_t)c->ndof));
    CK(c->bc_affected.alloc(c->N));
    CK(cudaMemset(c->bc_affected.p, 0, c->N));
    CK(c->bc_rows.alloc(c->N));
    CK(c->fixdof.alloc((size_t)c->ndof));
    CK(c->fixval.alloc((size_t)c->ndof));
    c->mf.affected = c->bc_affected.p;
    mf_prepare_geom(c->mf, c->E);
    return 0;
}

static int phase_begin(en234gpu_ctx* c) { CK(cudaSetDevice(c->device)); CK(cudaEventRecord(c->ev0, c->stream)); return 0; }
static int phase_end(en234gpu_ctx* c, double* ms) {
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float t = 0;
    CK(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    *ms = t;
    return 0;
}

// sum a per-node-owned norm^2 partial array into S->norm2, all-reduce across ranks, return sqrt
static TieTables tie_tables(en234gpu_ctx* c);
static int ties_after_assemble(en234gpu_ctx* c, double* nodal_force_norm);
static int ties_after_bc(en234gpu_ctx* c, double* unbalanced);
static int ties_before_solve(en234gpu_ctx* c);
static int ties_after_solve(en234gpu_ctx* c, double* sol, double* dl2);

static int finish_norm(en234gpu_ctx* c, const double* part, int n, double* out) {
    k_reduce_to<<<1, VEC_THREADS, 0, c->stream>>>(part, n, &c->S.p->norm2);
    c->st.kernel_launches++;
    if (c->comm.active && comm_allreduce_sum(c->comm, &c->S.p->norm2, 1, c->stream)) return set_err(comm_error());
    if (read_scalars(c)) return 1;
    *out = std::sqrt(c->hS->norm2);
    return 0;
}

static int assemble_generic(en234gpu_ctx* c, double* nodal_force_norm, int* fail) {
    if (c->mode != EN234_SOLVE_PCG_BSR) return set_err("4-, 10- and 20-node elements: assembled operator only");
    if (c->comm.active) return set_err("4-, 10- and 20-node elements: one device only");
    const GenMesh M = gen_mesh_of(c);
    const long long nb = (long long)c->E * c->npe * c->npe, nr = (long long)c->E * c->npe;
    k_gen_blocks<<<(unsigned)std::min<long long>((nb + 127) / 128, 65535LL * 16), 128, 0, c->stream>>>(M, c->Ke.p, c->S.p);
    k_gen_residual<<<(unsigned)std::min<long long>((nr + 127) / 128, 65535LL * 16), 128, 0, c->stream>>>(M, c->ut.p, c->du.p, c->Re.p, c->S.p);
    GenGatherArgs G;
    G.M = M; G.Ke = c->Ke.p; G.Re = c->Re.p; G.val = c->val.p; G.rhs = c->rhs.p; G.rforce = c->rforce.p;
    G.partials = c->part.p; G.diag = c->diag.p; G.S = c->S.p;
    k_gen_gather<<<c->grid_gather, THREADS, 0, c->stream>>>(G);
    c->st.kernel_launches += 3;
    CK(cudaGetLastError());
    c->diag_valid = true;
    c->matrix_valid = true;
    c->rhs_valid = true;
    if (finish_norm(c, c->part.p, c->grid_gather, nodal_force_norm)) return 1;
    *fail = c->hS->nanflag ? 1 : 0;
    if (c->ties.n && !*fail && ties_after_assemble(c, nodal_force_norm)) return 1;
    return phase_end(c, &c->st.assemble_ms);
}

int en234gpu_assemble_dev(en234gpu_handle c, double* nodal_force_norm, int* fail) {
    if (phase_begin(c)) return 1;
    CK(cudaMemsetAsync(&c->S.p->nanflag, 0, sizeof(int), c->stream));
    if (c->npe != 8) return assemble_generic(c, nodal_force_norm, fail);
    const int with_matrix = c->mode == EN234_SOLVE_PCG_BSR ? 1 : 0;
    // partitioned runs add the (full-valued) element load vector after the halo sum of the partial residuals
    const double* felem = (c->have_felem && !c->comm.active) ? c->felem.p : nullptr;
    int npart_asm = 0;
    if (c->asm_rows && !c->any_c3d8) {
        AsmArgs A;
        A.M = mesh_of(c);
        A.ut = c->ut.p; A.du = c->du.p;
        A.felem = felem;
        A.val = c->val.p; A.rhs = c->rhs.p; A.rforce = c->rforce.p;
        A.partials = c->part.p; A.S = c->S.p;
        A.assemble_matrix = with_matrix;
        if (c->any_neo) k_assemble_rows<true><<<c->grid_asm, c->asm_threads, c->asm_smem, c->stream>>>(A);
        else k_assemble_rows<false><<<c->grid_asm, c->asm_threads, c->asm_smem, c->stream>>>(A);
        c->st.kernel_launches++;
        npart_asm = c->grid_asm;
    } else {
        const size_t need_ke = with_matrix ? (size_t)((c->E + 3) / 4) * 4 * 576 : 0, need_re = (size_t)c->E * 24;
        if (c->Ke.n < need_ke && c->Ke.alloc(need_ke) != cudaSuccess) {
            cudaGetLastError();
            return set_err("en234gpu_assemble: cannot allocate the element scratch (" + std::to_string(need_ke * 8 >> 20) +
                           " MiB); set EN234_ASM=rows for the single-kernel row-gather assembly");
        }
        if (c->Re.n < need_re) CK(c->Re.alloc(need_re));
        ElemArgs EA;
        EA.M = mesh_of(c);
        EA.Ke = c->Ke.p; EA.Re = c->Re.p; EA.S = c->S.p;
        const LoadSum ld{c->ut.p, c->du.p};
        if (c->any_c3d8) {
            C3d8Args CA;
            CA.M = EA.M; CA.ut = c->ut.p; CA.du = c->du.p; CA.sv0 = c->sv0.p; CA.sv1 = c->sv1.p;
            CA.Ke = with_matrix ? c->Ke.p : nullptr; CA.Re = c->Re.p; CA.S = c->S.p;
            static const bool tpe = getenv("EN234_C3D8_TPE") && atoi(getenv("EN234_C3D8_TPE")) == 1;   // A/B: thread per element
            if (tpe) k_element_c3d8<<<(c->E + 63) / 64, 64, 0, c->stream>>>(CA);
            else {
                // per device (several handles on different GPUs may live in one process): set on every launch, it is cheap
                cudaFuncSetAttribute(k_element_c3d8_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(C3D8_WARPS * sizeof(C3d8Warp)));
                const int per_cta = C3D8_WARPS * C3D8_GROUP;
                const int grid = std::min((c->E + per_cta - 1) / per_cta, 148 * 12);
                k_element_c3d8_warp<<<grid, 32 * C3D8_WARPS, C3D8_WARPS * sizeof(C3d8Warp), c->stream>>>(CA);
            }
        } else if (c->any_neo && with_matrix) k_element_blocks<true, true, LoadSum><<<c->grid_elem, c->elem_threads, c->elem_smem, c->stream>>>(EA, ld);
        else if (c->any_neo) k_element_blocks<true, false, LoadSum><<<c->grid_elem0, c->elem_threads, c->elem_smem, c->stream>>>(EA, ld);
        else if (with_matrix) k_element_blocks<false, true, LoadSum><<<c->grid_elem, c->elem_threads, c->elem_smem, c->stream>>>(EA, ld);
        else launch_element_forces(c->mf, EA, ld, c->grid_elem0, c->stream);
        GatherArgs G;
        G.M = EA.M; G.Ke = c->Ke.p; G.Re = c->Re.p; G.felem = felem;
        G.val = c->val.p; G.rhs = c->rhs.p; G.rforce = c->rforce.p; G.partials = c->part.p; G.S = c->S.p;
        G.assemble_matrix = with_matrix;
        if (with_matrix && c->diag.n < (size_t)c->ndof) CK(c->diag.alloc((size_t)c->ndof));
        G.diag = c->diag.p; G.inv = c->inv.p;
        if (with_matrix) {
            k_gather_rows<<<c->grid_gather, THREADS, 0, c->stream>>>(G);
            npart_asm = c->grid_gather;
        } else {
            npart_asm = std::min(MAX_PARTIALS, std::max(1, (c->N + VEC_THREADS - 1) / VEC_THREADS));
            k_gather_residual<<<npart_asm, VEC_THREADS, 0, c->stream>>>(G);
        }
        c->st.kernel_launches += 2;
    }
    CK(cudaGetLastError());
    c->diag_valid = with_matrix && !c->asm_rows;
    c->matrix_valid = c->mode == EN234_SOLVE_PCG_BSR;
    c->rhs_valid = true;
    const double* part = c->part.p;
    int npart = npart_asm;
    if (c->comm.active) {
        // interface nodes hold partial sums: halo-sum rhs (and rforce = -rhs), then an owned-only norm
        if (comm_halo_sum(c->comm, c->rhs.p, c->stream)) return set_err(comm_error());
        k_negate_norm<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, c->rhs.p, c->have_felem ? c->felem.p : nullptr,
                                                                   c->rforce.p, c->owned.p, c->part2.p);
        c->st.kernel_launches += c->comm.launches_per_halo + 1;
        part = c->part2.p; npart = c->grid_vec;
    }
    if (finish_norm(c, part, npart, nodal_force_norm)) return 1;
    int bad = c->hS->nanflag;
    if (c->comm.active && comm_host_max(c->comm, &bad, c->stream)) return set_err(comm_error());
    *fail = bad ? 1 : 0;
    if (c->ties.n && !bad && ties_after_assemble(c, nodal_force_norm)) return 1;
    return phase_end(c, &c->st.assemble_ms);
}

static int upload_fixed(en234gpu_ctx* c, int n_fixed, const int* fixed_dof, const double* fixed_value) {
    // a DOF listed twice keeps its last value (the reference applies fixdof in list order)
    std::vector<int> last(n_fixed);
    std::map<int, int> seen;
    std::vector<int> d;
    std::vector<double> v;
    for (int i = 0; i < n_fixed; ++i) {
        if (fixed_dof[i] < 0 || fixed_dof[i] >= c->ndof) return set_err("prescribed DOF index out of range");
        auto it = seen.find(fixed_dof[i]);
        if (it == seen.end()) { seen[fixed_dof[i]] = (int)d.size(); d.push_back(fixed_dof[i]); v.push_back(fixed_value[i]); }
        else v[it->second] = fixed_value[i];
    }
    c->n_fixed = (int)d.size();
    if (d != c->h_fixdof || !c->bc_rows_valid) {
        c->bc_rows_valid = true;
        // rows changed by fixdof: the prescribed nodes and every node coupled to one (the pattern is symmetric)
        std::vector<uint8_t> aff(c->N, 0);
        for (int dof : d) {
            const int n = dof / 3;
            for (int k = c->h_browptr[n]; k < c->h_browptr[n + 1]; ++k) aff[c->h_bcol[k]] = 1;
            aff[n] = 1;
        }
        std::vector<int> rows;
        for (int i = 0; i < c->N; ++i) if (aff[i]) rows.push_back(i);
        c->n_bc_rows = (int)rows.size();
        if (c->bc_affected.n == 0) CK(c->bc_affected.alloc(c->N));
        if (c->bc_rows.n < rows.size()) CK(c->bc_rows.alloc(std::max(rows.size(), (size_t)1)));
        CK(cudaMemcpyAsync(c->bc_affected.p, aff.data(), aff.size(), cudaMemcpyHostToDevice, c->stream));
        if (!rows.empty()) CK(cudaMemcpyAsync(c->bc_rows.p, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->h_fixdof = d;
        c->mf.affected = c->bc_affected.p;
    }
    if (c->fixdof.n < d.size()) { CK(c->fixdof.alloc(d.size())); CK(c->fixval.alloc(d.size())); }
    if (!d.empty()) {
        CK(cudaMemcpyAsync(c->fixdof.p, d.data(), d.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->fixval.p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

static int apply_bc_common(en234gpu_ctx* c, int values_are_corrections, double* unbalanced) {
    if (!c->rhs_valid) return set_err("apply_boundaryconditions called before assemble");
    CK(cudaMemsetAsync(c->cflag.p, 0, (size_t)c->ndof, c->stream));
    if (c->n_fixed > 0) {
        k_set_corrections<<<(c->n_fixed + 255) / 256, 256, 0, c->stream>>>(c->n_fixed, c->fixdof.p, c->fixval.p, c->ut.p,
                                                                         c->du.p, c->cflag.p, c->cval.p,
                                                                         values_are_corrections);
        c->st.kernel_launches++;
    }
    BcArgs B;
    B.N = c->N; B.browptr = c->browptr.p; B.bcol = c->bcol.p; B.val = c->val.p; B.rhs = c->rhs.p;
    B.fext = c->have_fext ? c->fext.p : nullptr;
    B.cflag = c->cflag.p; B.cval = c->cval.p; B.minv = c->minv.p; B.partials = c->part.p;
    const double* part = c->part.p;
    int npart = c->grid_bc;
    const bool subset = c->mode == EN234_SOLVE_PCG_BSR && c->diag_valid && c->bc_rows_valid;
    int grid_bc = c->grid_bc;
    if (subset) {
        B.rows = c->bc_rows.p; B.n_rows = c->n_bc_rows;
        grid_bc = std::max(1, std::min(c->grid_bc, (c->n_bc_rows + WARPS - 1) / WARPS));
    }
    if (c->mode == EN234_SOLVE_PCG_MATRIXFREE) {
        // matrix-free: K[:,C] c is one application of the unmasked operator to the vector of prescribed corrections;
        // the Jacobi diagonal is integrated element by element
        const uint8_t* owned = c->have_owned ? c->owned.p : nullptr;
        const int nb = (c->ndof + 255) / 256;
        k_mask_values<<<nb, 256, 0, c->stream>>>(c->ndof, c->cflag.p, c->cval.p, c->tmpvec.p);
        mf_apply(c->mf, mesh_of(c), c->ndof, nullptr, nullptr, c->tmpvec.p, c->tmpvec2.p, c->part3.p, nullptr, c->stream,
                 &c->st.kernel_launches);
        mf_diag(c->mf, mesh_of(c), c->ndof, c->tmpvec.p, c->Ap.p, c->stream, &c->st.kernel_launches);
        if (c->comm.active) {
            if (comm_halo_sum(c->comm, c->tmpvec2.p, c->stream)) return set_err(comm_error());
            if (comm_halo_sum(c->comm, c->Ap.p, c->stream)) return set_err(comm_error());
        }
        k_fix_diag<<<nb, 256, 0, c->stream>>>(c->ndof, c->cflag.p, nullptr, c->Ap.p);
        k_bc_finalize<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, c->rhs.p, B.fext, c->cflag.p, c->cval.p,
                                                                   c->tmpvec2.p, c->Ap.p, c->minv.p, owned, c->part2.p);
        c->st.kernel_launches += 3;
        part = c->part2.p; npart = c->grid_vec;
    } else if (c->comm.active) {
        // the Jacobi diagonal and the rhs corrections of interface rows are partial sums on each rank:
        // apply to local rows, then halo-sum what is additive (see k_apply_bc_partitioned)
        if (bc_partitioned(c->comm, B, grid_bc, c->grid_vec, c->ndof, c->owned.p, c->tmpvec.p, c->tmpvec2.p, c->diag.p,
                           c->part2.p, c->stream, &c->st.kernel_launches)) return set_err(comm_error());
        part = c->part2.p; npart = c->grid_vec;
    } else if (subset) {
        // prescribed rows and their neighbours through the matrix; every other row from the assembled diagonal
        B.partials = c->part.p + c->grid_vec;
        k_bc_rest<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, c->bc_affected.p, c->rhs.p, B.fext, c->diag.p,
                                                               c->minv.p, c->part.p);
        if (c->n_bc_rows > 0) k_apply_bc<<<grid_bc, THREADS, 0, c->stream>>>(B);
        c->st.kernel_launches += c->n_bc_rows > 0 ? 2 : 1;
        npart = c->grid_vec + (c->n_bc_rows > 0 ? grid_bc : 0);
    } else {
        k_apply_bc<<<c->grid_bc, THREADS, 0, c->stream>>>(B);
        c->st.kernel_launches++;
    }
    CK(cudaGetLastError());
    if (finish_norm(c, part, npart, unbalanced)) return 1;
    if (c->ties.n && ties_after_bc(c, unbalanced)) return 1;
    return 0;
}

int en234gpu_apply_boundaryconditions_dev(en234gpu_handle c, double* unbalanced_force_norm) {
    if (phase_begin(c)) return 1;
    if (apply_bc_common(c, 0, unbalanced_force_norm)) return 1;
    return phase_end(c, &c->st.bc_ms);
}

// ---- PCG ------------------------------------------------------------------------------------------------
static int operator_partials(en234gpu_ctx* c, int method);
static CgVecArgs vec_args(en234gpu_ctx* c, int method, int parity) {
    CgVecArgs V{};
    V.n = c->ndof; V.rhs = c->rhs.p; V.minv = c->minv.p;
    V.owned = c->have_owned ? c->owned.p : nullptr;
    V.sol = c->sol.p; V.r = c->r.p; V.p = c->p.p; V.Ap = c->Ap.p;
    V.part_pAp = c->part.p;
    const bool fused = c->comm.active && c->comm.p2p && c->comm.fused && c->comm.fused_loop;
    V.n_part_pAp = (c->comm.active && !fused) ? 1 : operator_partials(c, method);
    V.part_a = c->part2.p; V.part_b = c->part3.p;
    V.n_part = (c->comm.active && !fused) ? 1 : c->grid_vec;
    V.S = c->S.p;
    V.parity = parity;
    if (c->comm.active && c->comm.p2p && c->comm.fused && c->comm.fused_loop) {
        V.F = c->comm.F;
        V.F.halo_in_spmv = method == EN234_SOLVE_PCG_BSR ? 1 : 0;
        V.F.pAp_pushed = method == EN234_SOLVE_PCG_BSR ? 1 : 0;
    }
    return V;
}

// EN234_P2P_LL=0: the all-reduce kernels of the PCG loop in their first form (value, fence, flag) instead of self-validating packets
static bool p2p_ll() {
    static const bool on = !(getenv("EN234_P2P_LL") && atoi(getenv("EN234_P2P_LL")) == 0);
    return on;
}

// EN234_P2P_TAIL: 1 = the producers' last CTAs push the rank's sums, single-warp wait kernels consume them; 2 = the last
// CTAs are producer AND consumer (push, wait for the peers, add): no collective kernel for the two all-reduces at all
static int p2p_tail_mode() {
    static const int mode = getenv("EN234_P2P_TAIL") ? atoi(getenv("EN234_P2P_TAIL")) : 0;
    return (mode == 1 || mode == 2) ? mode : 0;
}

// index of the symmetric-storage product (k_spmv_sym): built on first use
static bool spmv_sym_ready(en234gpu_ctx* c) {
    if (c->spmv_sym != 0) return c->spmv_sym > 0;
    c->spmv_sym = -1;
    const bool want = getenv("EN234_SPMV_SYM") && atoi(getenv("EN234_SPMV_SYM")) == 1;      // read once per handle
    if (!want || c->any_c3d8 || c->max_row_blocks > 32 || c->h_browptr.empty() || 9ULL * (unsigned long long)c->nnzb >= 0xffffffffULL) return false;
    const std::vector<int>& rp = c->h_browptr;
    const std::vector<int>& bc = c->h_bcol;
    std::vector<unsigned> base(bc.size(), 0xffffffffu);
    std::vector<uint8_t> len(bc.size(), 0);
    for (int i = 0; i < c->N; ++i)
        for (int k = rp[i]; k < rp[i + 1] && bc[k] < i; ++k) {
            const int j = bc[k];
            const int* b0 = bc.data() + rp[j];
            const int* b1 = bc.data() + rp[j + 1];
            const int* it = std::lower_bound(b0, b1, i);
            if (it == b1 || *it != i) return false;               // pattern not symmetric: keep the full product
            base[k] = (unsigned)(9ULL * (unsigned long long)rp[j] + 3ULL * (unsigned long long)(it - b0));
            len[k] = (uint8_t)(3 * (rp[j + 1] - rp[j]));
        }
    if (c->sym_base.alloc(base.size()) != cudaSuccess || c->sym_len.alloc(len.size()) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaMemcpy(c->sym_base.p, base.data(), base.size() * sizeof(unsigned), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->sym_len.p, len.data(), len.size(), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    c->spmv_sym = 1;
    return true;
}

// y = A x on block rows [row0, row1) (assembled operator); partials are written at part[part_off ...]
// CTA slots the interior-row SpMV of a partitioned run leaves free (EN234_P2P_RESERVE, default 32): its persistent grid
// otherwise fills every SM until it ends, so the halo push / wait kernels of the side stream (32-64 CTAs) start only after
// it.  Measured on 2 B200 with the packet all-reduce (64^3 per GPU): 0 -> 116.4, 16 -> 112.6, 32 -> 113.0, 64 -> 113.5,
// 96 -> 117.0 us per PCG iteration.
static int p2p_reserve() {
    static const int n = getenv("EN234_P2P_RESERVE") ? std::max(0, atoi(getenv("EN234_P2P_RESERVE"))) : 32;
    return n;
}

static int launch_spmv_rows(en234gpu_ctx* c, const double* x, double* y, int row0, int row1, int part_off, int check_stop,
                            int* n_part, int tail_total = 0, cudaStream_t stream = nullptr, bool pdl = false, int reserve = 0) {
    if (!stream) stream = c->stream;
    const int need = (row1 - row0 + WARPS - 1) / WARPS;
    const int grid = std::max(1, std::min(need, std::max(1, c->grid_spmv - reserve)));
    SpmvArgs A;
    A.row0 = row0; A.N = row1; A.browptr = c->browptr.p; A.bcol = c->bcol.p; A.val = c->val.p; A.x = x; A.y = y;
    A.partials = c->part.p + part_off; A.S = c->S.p; A.check_stop = check_stop;
    if (tail_total > 0) {
        A.tail = (ArenaHdr*)c->comm.arena; A.tail_partials = c->part.p; A.tail_total = tail_total;
        if (p2p_tail_mode() == 2) A.tail_out = c->part.p;
    }
    if (c->mode == EN234_SOLVE_PCG_BSR && check_stop && spmv_sym_ready(c)) k_spmv_sym<4><<<grid, THREADS, 0, stream>>>(A, c->sym_base.p, c->sym_len.p);
    else if (c->spmv_occ == 5) k_spmv<5><<<grid, THREADS, 0, stream>>>(A);
    else if (pdl && c->max_row_blocks <= 32) launch_pdl(k_spmv<4, false, false>, grid, THREADS, stream, A, FusedComm{});
    else if (pdl) launch_pdl(k_spmv<4, true, false>, grid, THREADS, stream, A, FusedComm{});
    else if (c->max_row_blocks <= 32) k_spmv<4, false><<<grid, THREADS, 0, stream>>>(A);
    else k_spmv<4><<<grid, THREADS, 0, stream>>>(A);
    c->st.kernel_launches++;
    *n_part = grid;
    return 0;
}

static int launch_operator(en234gpu_ctx* c, int method, const double* x, double* y, int check_stop) {
    if (method == EN234_SOLVE_PCG_MATRIXFREE) {
        mf_apply(c->mf, mesh_of(c), c->ndof, c->cflag.p, c->have_owned ? c->owned.p : nullptr, x, y, c->part.p,
                 check_stop ? c->S.p : nullptr, c->stream, &c->st.kernel_launches);
    } else {
        int np = 0;
        launch_spmv_rows(c, x, y, 0, c->N, 0, check_stop, &np, 0, nullptr, check_stop != 0);
    }
    return 0;
}

// interface rows of the assembled operator on a partitioned run: the fused SpMV variant stores them into the neighbours'
// arenas as it goes and raises the flags; partials go to part[0 ...)
static int launch_spmv_interface(en234gpu_ctx* c, const double* x, double* y, cudaStream_t stream, int tail_total) {
    const int n_if = c->comm.n_interface;
    const int need = (n_if + WARPS - 1) / WARPS;
    const int grid = std::max(1, std::min(need, c->grid_spmv));
    SpmvArgs A;
    A.row0 = 0; A.N = n_if; A.browptr = c->browptr.p; A.bcol = c->bcol.p; A.val = c->val.p; A.x = x; A.y = y;
    A.partials = c->part.p; A.S = c->S.p; A.check_stop = 1;
    A.tail = (ArenaHdr*)c->comm.arena; A.tail_partials = c->part.p; A.tail_total = tail_total;
    FusedComm F = c->comm.F;
    F.halo_in_spmv = 1;
    if (c->max_row_blocks <= 32) k_spmv<4, false, true><<<grid, THREADS, 0, stream>>>(A, F);
    else k_spmv<4, true, true><<<grid, THREADS, 0, stream>>>(A, F);
    c->st.kernel_launches++;
    return grid;
}

// peer-memory halo PUSH only (the consumer k_cg_update adds what arrived), forked to the side stream
static void p2p_halo_push_fork(en234gpu_ctx* c, double* vec) {
    Comm& cm = c->comm;
    cudaEventRecord(cm.ev_fork, c->stream);
    cudaStreamWaitEvent(cm.side, cm.ev_fork, 0);
    if (cm.n_nbr > 0) {
        int mx = 0;
        for (int k = 0; k < cm.n_nbr; ++k) mx = std::max(mx, 3 * (cm.shared_ptr[k + 1] - cm.shared_ptr[k]));
        const int nblk = std::max(1, std::min(32, (mx + 255) / 256));
        k_halo_push<<<dim3(nblk, cm.n_nbr), 256, 0, cm.side>>>(cm.peers, cm.halo, vec, c->S.p, 1);
        c->st.kernel_launches += 1;
    }
    cudaEventRecord(cm.ev_join, cm.side);
}

// number of p.Ap partials the operator leaves in part[]
static int operator_partials(en234gpu_ctx* c, int method) {
    if (method == EN234_SOLVE_PCG_MATRIXFREE) return c->mf.n_partials;
    if (c->comm.active && c->comm.p2p && c->comm.fused && c->comm.fused_loop) {
        // two launches: interface rows, interior rows
        const int n_if = c->comm.n_interface;
        const int n1 = n_if > 0 ? std::max(1, std::min((n_if + WARPS - 1) / WARPS, c->grid_spmv)) : 0;
        const int n2 = n_if < c->N ? std::max(1, std::min((c->N - n_if + WARPS - 1) / WARPS, c->grid_spmv)) : 0;
        return n1 + n2;
    }
    return std::max(1, std::min((c->N + WARPS - 1) / WARPS, c->grid_spmv));
}

// peer-memory halo sum of `vec` on the side stream, forked from / joined to the main stream by events
static void p2p_halo_fork(en234gpu_ctx* c, double* vec, int check_stop) {
    Comm& cm = c->comm;
    cudaEventRecord(cm.ev_fork, c->stream);
    cudaStreamWaitEvent(cm.side, cm.ev_fork, 0);
    if (cm.n_nbr > 0) {
        int mx = 0;
        for (int k = 0; k < cm.n_nbr; ++k) mx = std::max(mx, 3 * (cm.shared_ptr[k + 1] - cm.shared_ptr[k]));
        const int nblk = std::max(1, std::min(32, (mx + 255) / 256));
        k_halo_push<<<dim3(nblk, cm.n_nbr), 256, 0, cm.side>>>(cm.peers, cm.halo, vec, c->S.p, check_stop);
        if (cm.fused) {
            // one launch: wait for all neighbours, add own and received values in ascending rank order
            k_halo_wait_total<<<std::max(1, std::min(64, (3 * cm.n_interface + 255) / 256)), 256, 0, cm.side>>>(cm.F, vec, c->S.p, check_stop);
            c->st.kernel_launches += 2;
        } else {
            for (int k = 0; k < cm.n_nbr; ++k)
                k_halo_wait_add<<<nblk, 256, 0, cm.side>>>(cm.peers, cm.halo, k, k == cm.n_nbr - 1 ? 1 : 0, vec, c->S.p, check_stop);
            c->st.kernel_launches += 1 + cm.n_nbr;
        }
    }
    cudaEventRecord(cm.ev_join, cm.side);
}
static void p2p_halo_join(en234gpu_ctx* c) { cudaStreamWaitEvent(c->stream, c->comm.ev_join, 0); }
// the same exchange for a caller that has already forked the side stream; phase 1 = push only, 2 = wait + add only (the
// single-stream form puts the interior rows between the two), 3 = both on the side stream + the join event
static void p2p_halo_exchange_on_side(en234gpu_ctx* c, double* vec, int check_stop, int phase = 3, cudaStream_t stream = nullptr) {
    Comm& cm = c->comm;
    if (!stream) stream = cm.side;
    if (cm.n_nbr > 0) {
        int mx = 0;
        for (int k = 0; k < cm.n_nbr; ++k) mx = std::max(mx, 3 * (cm.shared_ptr[k + 1] - cm.shared_ptr[k]));
        const int nblk = std::max(1, std::min(32, (mx + 255) / 256));
        if (phase & 1) {
            k_halo_push<<<dim3(nblk, cm.n_nbr), 256, 0, stream>>>(cm.peers, cm.halo, vec, c->S.p, check_stop);
            c->st.kernel_launches += 1;
        }
        if ((phase & 2) && cm.fused) {
            // one launch: wait for all neighbours, add own and received values in ascending rank order
            k_halo_wait_total<<<std::max(1, std::min(64, (3 * cm.n_interface + 255) / 256)), 256, 0, stream>>>(cm.F, vec, c->S.p, check_stop);
            c->st.kernel_launches += 1;
        } else if (phase & 2) {
            for (int k = 0; k < cm.n_nbr; ++k)
                k_halo_wait_add<<<nblk, 256, 0, stream>>>(cm.peers, cm.halo, k, k == cm.n_nbr - 1 ? 1 : 0, vec, c->S.p, check_stop);
            c->st.kernel_launches += cm.n_nbr;
        }
    }
    if (phase == 3) cudaEventRecord(cm.ev_join, cm.side);
}

// one PCG iteration on the stream (captured into the graph)
static int enqueue_iteration(en234gpu_ctx* c, int method, int parity) {
    Comm& cm = c->comm;
    t_pdl_auto = cm.active && cm.p2p;
    if (cm.active && cm.p2p && cm.fused && cm.fused_loop) {
        // collectives folded into the kernels: no collective has a launch of its own
        if (method == EN234_SOLVE_PCG_BSR) {
            // interface rows first (they push themselves to the neighbours), then the interior rows with the lean kernel
            // (the few interface rows run on the side stream, concurrently with the interior rows)
            // The CTA that finishes last over both launches reduces the p.Ap partials and pushes the rank's sum.
            int n1 = 0, n2 = 0;
            const int total = operator_partials(c, method);
            if (cm.n_interface > 0) {
                cudaEventRecord(cm.ev_fork, c->stream);
                cudaStreamWaitEvent(cm.side, cm.ev_fork, 0);
                n1 = launch_spmv_interface(c, c->p.p, c->Ap.p, cm.side, total);
                cudaEventRecord(cm.ev_join, cm.side);
            }
            if (cm.n_interface < c->N) launch_spmv_rows(c, c->p.p, c->Ap.p, cm.n_interface, c->N, n1, 1, &n2, total);
            if (cm.n_interface > 0) cudaStreamWaitEvent(c->stream, cm.ev_join, 0);
        } else {
            const uint8_t* owned = c->have_owned ? c->owned.p : nullptr;
            if (c->mf.congruent) {
                mf_stencil(c->mf, mesh_of(c), c->cflag.p, owned, c->p.p, c->Ap.p, c->part.p, c->S.p, c->stream, &c->st.kernel_launches, 0);
                p2p_halo_push_fork(c, c->Ap.p);
                mf_stencil(c->mf, mesh_of(c), c->cflag.p, owned, c->p.p, c->Ap.p, c->part.p, c->S.p, c->stream, &c->st.kernel_launches, 1);
            } else {
                launch_operator(c, method, c->p.p, c->Ap.p, 1);
                p2p_halo_push_fork(c, c->Ap.p);
            }
            p2p_halo_join(c);
        }
        CgVecArgs V = vec_args(c, method, parity);
        k_cg_update<true><<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(V);
        k_cg_pupdate<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(V);
        c->st.kernel_launches += 2;
        return 0;
    }
    if (cm.active && cm.p2p) {
        // interface rows first; their halo exchange (peer-memory stores + flags, side stream) overlaps the interior
        // rows; p.Ap is the all-reduced sum of local partial products over ALL local rows, which needs no halo first
        // EN234_P2P_TAIL=1 (A/B; measured 3 us per iteration slower on 2 B200): the producers' last CTAs reduce their partials
        // and push the rank's sums (the stores cross NVLink during the kernel boundary), the collective kernels in between
        // only wait and add.  Default: single-CTA all-reduce kernels (reduce, push, wait, add).
        const bool tail_push = p2p_tail_mode() != 0, tail_finish = p2p_tail_mode() == 2;
        bool spmv_pushed = false;
        int np = 0;
        static const bool serial = getenv("EN234_P2P_SERIAL") && atoi(getenv("EN234_P2P_SERIAL")) == 1;
        bool joined = false;
        if (serial && method == EN234_SOLVE_PCG_BSR && cm.n_interface > 0 && cm.n_interface < c->N) {
            // ONE stream: interface rows -> push -> interior rows -> wait + add.  The pushed values cross NVLink while the
            // interior rows run; no fork / join edges between graph branches.
            int n1 = 0, n2 = 0;
            launch_spmv_rows(c, c->p.p, c->Ap.p, 0, cm.n_interface, 0, 1, &n1);
            p2p_halo_exchange_on_side(c, c->Ap.p, 1, 1, c->stream);
            launch_spmv_rows(c, c->p.p, c->Ap.p, cm.n_interface, c->N, n1, 1, &n2);
            p2p_halo_exchange_on_side(c, c->Ap.p, 1, 2, c->stream);
            np = n1 + n2;
            joined = true;
        } else if (method == EN234_SOLVE_PCG_BSR && cm.n_interface > 0 && cm.n_interface < c->N) {
            // side stream: the few interface rows, then their exchange (push, wait + add); main stream: the interior rows
            // at the same time.  The all-reduce needs the partials of both launches (ev_if), k_cg_update the whole side
            // stream (ev_join).
            int n1 = 0, n2 = 0;
            cudaEventRecord(cm.ev_fork, c->stream);
            cudaStreamWaitEvent(cm.side, cm.ev_fork, 0);
            const int g1 = std::max(1, std::min((cm.n_interface + WARPS - 1) / WARPS, c->grid_spmv));
            const int g2 = std::max(1, std::min((c->N - cm.n_interface + WARPS - 1) / WARPS, std::max(1, c->grid_spmv - p2p_reserve())));
            const int tail_total = tail_push ? g1 + g2 : 0;
            launch_spmv_rows(c, c->p.p, c->Ap.p, 0, cm.n_interface, 0, 1, &n1, tail_total, cm.side);
            cudaEventRecord(cm.ev_if, cm.side);
            p2p_halo_exchange_on_side(c, c->Ap.p, 1);
            launch_spmv_rows(c, c->p.p, c->Ap.p, cm.n_interface, c->N, n1, 1, &n2, tail_total, nullptr, use_pdl_interior_spmv(), p2p_reserve());
            cudaStreamWaitEvent(c->stream, cm.ev_if, 0);
            np = n1 + n2;
            spmv_pushed = tail_push;
        } else if (method == EN234_SOLVE_PCG_MATRIXFREE && c->mf.congruent) {
            // congruent-mesh operator: the non-dominant rows (all interface rows among them) first, their halo
            // exchange overlaps the dominant rows
            const uint8_t* owned = c->have_owned ? c->owned.p : nullptr;
            mf_stencil(c->mf, mesh_of(c), c->cflag.p, owned, c->p.p, c->Ap.p, c->part.p, c->S.p, c->stream, &c->st.kernel_launches, 0);
            p2p_halo_fork(c, c->Ap.p, 1);
            mf_stencil(c->mf, mesh_of(c), c->cflag.p, owned, c->p.p, c->Ap.p, c->part.p, c->S.p, c->stream, &c->st.kernel_launches, 1);
            np = operator_partials(c, method);
        } else {
            launch_operator(c, method, c->p.p, c->Ap.p, 1);
            np = operator_partials(c, method);
            p2p_halo_fork(c, c->Ap.p, 1);
        }
        if (spmv_pushed && tail_finish) { /* the last SpMV CTA left the total in part[0] */ }
        else if (spmv_pushed) k_p2p_wait<<<1, 32, 0, c->stream>>>(cm.peers, c->part.p, nullptr, c->S.p);
        else launch_pdl(p2p_ll() ? k_p2p_allreduce : k_p2p_allreduce_fence, 1, VEC_THREADS, c->stream, cm.peers, (const double*)c->part.p, np,
                        (const double*)nullptr, 0, c->part.p, (double*)nullptr, (const CgScalars*)c->S.p, 1);
        if (!joined) p2p_halo_join(c);
        CgVecArgs V = vec_args(c, method, parity);
        if (tail_push) { V.tail = (ArenaHdr*)cm.arena; V.tail_finish = tail_finish ? 1 : 0; }
        launch_pdl(k_cg_update<false>, c->grid_vec, VEC_THREADS, c->stream, V);
        if (tail_finish) { /* k_cg_update's last CTA left the totals in part2[0] / part3[0] */ }
        else if (tail_push) k_p2p_wait<<<1, 32, 0, c->stream>>>(cm.peers, c->part2.p, c->part3.p, c->S.p);
        else launch_pdl(p2p_ll() ? k_p2p_allreduce : k_p2p_allreduce_fence, 1, VEC_THREADS, c->stream, cm.peers, (const double*)c->part2.p,
                        c->grid_vec, (const double*)c->part3.p, c->grid_vec, c->part2.p, c->part3.p, (const CgScalars*)c->S.p, 1);
        launch_pdl(k_cg_pupdate, c->grid_vec, VEC_THREADS, c->stream, V);
        c->st.kernel_launches += tail_finish ? (spmv_pushed ? 2 : 3) : 4;
        return 0;
    }
    if (c->ties.n) {
        k_tie_expand<<<std::max(1, (c->ties.n_tdof + 255) / 256), 256, 0, c->stream>>>(tie_tables(c), c->p.p, c->S.p);
        c->st.kernel_launches++;
    }
    launch_operator(c, method, c->p.p, c->Ap.p, 1);
    if (c->ties.n) {
        k_tie_reduce<<<std::max(1, (c->ties.n_tdof + 255) / 256), 256, 0, c->stream>>>(tie_tables(c), c->Ap.p, c->p.p, c->S.p);
        c->st.kernel_launches++;
    }
    if (cm.active) {
        // NCCL path (EN234_COMM=nccl or no peer arena)
        const int np = operator_partials(c, method);
        k_reduce_to<<<1, VEC_THREADS, 0, c->stream>>>(c->part.p, np, c->part.p);
        if (comm_allreduce_sum(c->comm, c->part.p, 1, c->stream)) return 1;
        if (comm_halo_sum(c->comm, c->Ap.p, c->stream)) return 1;
        c->st.kernel_launches += 1 + c->comm.launches_per_halo;
    }
    CgVecArgs V = vec_args(c, method, parity);
    launch_pdl(k_cg_update<false>, c->grid_vec, VEC_THREADS, c->stream, V);
    c->st.kernel_launches++;
    if (cm.active) {
        k_reduce2_to<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, c->part3.p, c->grid_vec, c->part2.p, c->part3.p);
        // pack {rr, rz} in one 2-double message: part2[0], part2[1]
        k_pack2<<<1, 1, 0, c->stream>>>(c->part2.p, c->part3.p);
        if (comm_allreduce_sum(c->comm, c->part2.p, 2, c->stream)) return 1;
        k_unpack2<<<1, 1, 0, c->stream>>>(c->part2.p, c->part3.p);
        c->st.kernel_launches += 3;
    }
    launch_pdl(k_cg_pupdate, c->grid_vec, VEC_THREADS, c->stream, V);
    c->st.kernel_launches++;
    return 0;
}

static int build_graph(en234gpu_ctx* c, int method) {
    cudaGraph_t g = nullptr;
    const long long before = c->st.kernel_launches;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    for (int it = 1; it <= CG_CHUNK && !rc; ++it) rc = enqueue_iteration(c, method, it & 1);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->graph_launches[method] = c->st.kernel_launches - before;
    c->st.kernel_launches = before;          // capture does not launch
    if (rc) return set_err(std::string("graph capture: ") + comm_error());
    if (e != cudaSuccess) return set_err(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    CK(cudaGraphInstantiate(&c->graph[method], g, 0));
    CK(cudaGraphDestroy(g));
    return 0;
}

// ---- BiCGSTAB (unsymmetric stiffness) ---------------------------------------------------------------------
constexpr int BI_CHUNK = 8;     // BiCGSTAB iterations per CUDA-graph launch

// partitioned runs: the two partial arrays become global sums in partA[0], partB[0] (the single-CTA consumers then read
// one entry); returns the number of partials the consumers read
static int bicg_allreduce2(en234gpu_ctx* c, double* partA, double* partB, int n) {
    if (!c->comm.active) return n;
    if (c->comm.p2p) {
        // peer-memory arenas: one single-CTA kernel (reduce, store to every rank, wait, add in rank order)
        k_p2p_allreduce<<<1, VEC_THREADS, 0, c->stream>>>(c->comm.peers, partA, n, partB, n, partA, partB, c->S.p, 0);
        c->st.kernel_launches += 1;
        return 1;
    }
    k_reduce2_to<<<1, VEC_THREADS, 0, c->stream>>>(partA, partB, n, partA, partB);
    k_pack2<<<1, 1, 0, c->stream>>>(partA, partB);
    if (comm_allreduce_sum(c->comm, partA, 2, c->stream)) return -1;
    k_unpack2<<<1, 1, 0, c->stream>>>(partA, partB);
    c->st.kernel_launches += 3;
    return 1;
}
// y = A x on all local rows, interface rows halo-summed on partitioned runs
static int bicg_operator(en234gpu_ctx* c, const double* x, double* y) {
    int np = 0;
    const int gt = std::max(1, (c->ties.n_tdof + 255) / 256);
    // tie constraints: the operator of the reduced system, T^T A T (x has zeros on the members of every tied group)
    if (c->ties.n) k_tie_expand<<<gt, 256, 0, c->stream>>>(tie_tables(c), const_cast<double*>(x), nullptr);
    launch_spmv_rows(c, x, y, 0, c->N, 0, 0, &np);
    if (c->ties.n) {
        k_tie_reduce<<<gt, 256, 0, c->stream>>>(tie_tables(c), y, const_cast<double*>(x), nullptr);
        c->st.kernel_launches += 2;
    }
    if (c->comm.active) {
        if (comm_halo_sum(c->comm, y, c->stream)) return 1;
        c->st.kernel_launches += c->comm.launches_per_halo;
    }
    return 0;
}

static int bicg_enqueue_iteration(en234gpu_ctx* c) {
    const int n = c->ndof, gv = c->grid_vec, nb = (n + 255) / 256;
    double *x = c->sol.p, *r = c->r.p, *rhat = c->tmpvec.p, *p = c->p.p, *v = c->Ap.p, *y = c->tmpvec2.p, *z = c->bi_z.p, *t = c->bi_t.p;
    const uint8_t* owned = c->have_owned ? c->owned.p : nullptr;
    BiScalars* S = c->bi_S.p;
    k_bi_p<<<nb, 256, 0, c->stream>>>(n, S, r, v, c->minv.p, p, y);
    if (bicg_operator(c, y, v)) return 1;
    k_bi_dot2<<<gv, VEC_THREADS, 0, c->stream>>>(n, S, rhat, v, rhat, v, c->part2.p, c->part3.p, owned);
    int np = bicg_allreduce2(c, c->part2.p, c->part3.p, gv);
    if (np < 0) return 1;
    k_bi_alpha<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, np, S);
    k_bi_s<<<nb, 256, 0, c->stream>>>(n, S, v, c->minv.p, r, z);
    if (bicg_operator(c, z, t)) return 1;
    k_bi_dot2<<<gv, VEC_THREADS, 0, c->stream>>>(n, S, t, r, t, t, c->part2.p, c->part3.p, owned);
    np = bicg_allreduce2(c, c->part2.p, c->part3.p, gv);
    if (np < 0) return 1;
    k_bi_omega<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, c->part3.p, np, S);
    k_bi_x<<<gv, VEC_THREADS, 0, c->stream>>>(n, S, y, z, t, rhat, x, r, c->part2.p, c->part3.p, owned);
    np = bicg_allreduce2(c, c->part2.p, c->part3.p, gv);
    if (np < 0) return 1;
    k_bi_finish<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, c->part3.p, np, S);
    c->st.kernel_launches += 8;
    return 0;
}

// One BiCGSTAB run for A d = b from d = 0 (recurrence residual driven to (r.r)/(b.b) <= tol), result in c->sol; returns
// the iteration count, -1 on breakdown / iteration cap, -2 on a CUDA error
static int bicgstab_run(en234gpu_ctx* c, const double* b, double tol, int maxit) {
    const int n = c->ndof, gv = c->grid_vec;
    const uint8_t* owned = c->have_owned ? c->owned.p : nullptr;
    if (!c->bi_graph) {
        cudaGraph_t g = nullptr;
        const long long before = c->st.kernel_launches;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return -2;
        int rc = 0;
        for (int it = 0; it < BI_CHUNK && !rc; ++it) rc = bicg_enqueue_iteration(c);
        if (cudaStreamEndCapture(c->stream, &g) != cudaSuccess || rc) return -2;
        c->bi_graph_launches = c->st.kernel_launches - before;
        c->st.kernel_launches = before;
        if (cudaGraphInstantiate(&c->bi_graph, g, 0) != cudaSuccess) return -2;
        cudaGraphDestroy(g);
    }
    k_bi_init<<<gv, VEC_THREADS, 0, c->stream>>>(n, b, c->sol.p, c->r.p, c->tmpvec.p, c->p.p, c->Ap.p, c->part2.p, owned);
    int np = gv;
    if (c->comm.active) {
        k_reduce_to<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, gv, c->part2.p);
        if (comm_allreduce_sum(c->comm, c->part2.p, 1, c->stream)) return -2;
        c->st.kernel_launches += 1;
        np = 1;
    }
    k_bi_init_finish<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, np, c->bi_S.p, tol, maxit);
    c->st.kernel_launches += 2;
    for (;;) {
        if (cudaMemcpyAsync(c->h_bi_S, c->bi_S.p, sizeof(BiScalars), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -2;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -2;
        if (c->h_bi_S->stop) break;
        if (cudaGraphLaunch(c->bi_graph, c->stream) != cudaSuccess) return -2;
        c->st.kernel_launches += c->bi_graph_launches;
    }
    return c->h_bi_S->fail ? -1 : c->h_bi_S->iter;
}

// A x = rhs to a TRUE residual (b - A x).(b - A x)/(b.b) <= tol: BiCGSTAB runs on the current true residual, the
// corrections are accumulated (iterative refinement).  The Newton history of the reference's golden log needs solves of
// direct-solver quality (6 printed digits of a 1e-4 second correction), which a single recurrence-residual stop misses.
// Partitioned runs: halo sums of A y / A z on interface rows, owned-only dot products all-reduced over the ranks (every
// rank carries bitwise identical recurrence scalars, so all ranks stop in the same iteration).
static int bicgstab_solve(en234gpu_ctx* c, double tol, int maxit, double* correction_norm, int* iterations, int* fail) {
    if (c->mode != EN234_SOLVE_PCG_BSR || !c->matrix_valid) return set_err("BiCGSTAB needs the assembled matrix (solve called before assemble?)");
    if (phase_begin(c)) return 1;
    const int n = c->ndof, gv = c->grid_vec, nb = (n + 255) / 256;
    const uint8_t* owned = c->have_owned ? c->owned.p : nullptr;
    if (c->bi_z.n < (size_t)n) {
        CK(c->bi_z.alloc(n)); CK(c->bi_t.alloc(n)); CK(c->bi_xt.alloc(n)); CK(c->bi_b.alloc(n)); CK(c->bi_S.alloc(1));
        CK(cudaMallocHost((void**)&c->h_bi_S, sizeof(BiScalars)));
    }
    if (c->ties.n && ties_before_solve(c)) return 1;
    const double tol_run = std::max(tol, 1e-24);
    int total = 0;
    *fail = 0;
    double bb = 0.0, rr_prev = 0.0;
    for (int pass = 0; pass < 6; ++pass) {
        const double* b = pass == 0 ? c->rhs.p : c->bi_b.p;
        const int it = bicgstab_run(c, b, pass == 0 ? tol_run : 1e-12, std::max(1, maxit - total));
        if (it == -2) return set_err(std::string("BiCGSTAB: CUDA / communicator error ") + comm_error());
        if (it < 0) { *fail = 1; break; }
        total += it;
        k_bi_accumulate<<<nb, 256, 0, c->stream>>>(n, c->sol.p, c->bi_xt.p, pass == 0);
        if (bicg_operator(c, c->bi_xt.p, c->bi_t.p)) return set_err(comm_error());
        k_bi_residual<<<gv, VEC_THREADS, 0, c->stream>>>(n, c->rhs.p, c->bi_t.p, c->bi_b.p, c->part2.p, c->part3.p, owned);
        const int np = bicg_allreduce2(c, c->part2.p, c->part3.p, gv);
        if (np < 0) return set_err(comm_error());
        k_bi_reduce2<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, c->part3.p, np, &c->S.p->rz[0]);
        c->st.kernel_launches += 4;
        if (read_scalars(c)) return 1;
        const double rr = c->hS->rz[0];
        bb = c->hS->rz[1];
        c->st.last_cg_err = bb > 0.0 ? rr / bb : 0.0;
        if (!(bb > 0.0) || rr / bb <= tol) break;
        if (pass > 0 && rr > 0.25 * rr_prev) break;          // refinement has reached the rounding floor
        rr_prev = rr;
    }
    CK(cudaGetLastError());
    *iterations = total;
    c->st.last_cg_iterations = total;
    if (!*fail) {
        double dl2 = 0.0;
        if (c->ties.n && ties_after_solve(c, c->bi_xt.p, &dl2)) return 1;
        k_add_solution<<<gv, VEC_THREADS, 0, c->stream>>>(n, c->bi_xt.p, c->du.p, owned, c->part2.p);
        c->st.kernel_launches++;
        if (finish_norm(c, c->part2.p, gv, correction_norm)) return 1;
        if (c->ties.n) *correction_norm = std::sqrt(*correction_norm * *correction_norm + dl2);
    } else *correction_norm = 0.0;
    return phase_end(c, &c->st.solve_ms);
}

// PCG on (val | matrix-free operator, rhs, minv) -> sol: initialisation, then chunks of CG_CHUNK captured iterations; the
// stop flag is read back after each chunk (iterations after convergence inside a chunk return immediately).  Leaves the
// final scalars in c->hS.
static int pcg_run(en234gpu_ctx* c, int method, double cg_tolerance, int max_cg_iterations) {
    if (method == EN234_SOLVE_PCG_BSR) spmv_sym_ready(c);       // allocates on first use: must not happen inside the capture
    if (!c->graph[method] && build_graph(c, method)) return 1;
    // tolerance / iteration cap are run-time values in device memory (compile-time in modules/Stiffness.f90:35-37)
    CgScalars init{};
    init.tol = cg_tolerance; init.maxit = max_cg_iterations;
    CK(cudaMemcpyAsync(c->S.p, &init, sizeof(CgScalars), cudaMemcpyHostToDevice, c->stream));
    CgVecArgs V = vec_args(c, method, 1);
    V.n_part = c->grid_vec;
    k_cg_init<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(V);
    c->st.kernel_launches++;
    if (c->comm.active && c->comm.p2p) {
        k_p2p_allreduce<<<1, VEC_THREADS, 0, c->stream>>>(c->comm.peers, c->part2.p, c->grid_vec, c->part3.p, c->grid_vec,
                                                          c->part2.p, c->part3.p, c->S.p, 0);
        V.n_part = 1;
        c->st.kernel_launches += 1;
    } else if (c->comm.active) {
        k_reduce2_to<<<1, VEC_THREADS, 0, c->stream>>>(c->part2.p, c->part3.p, c->grid_vec, c->part2.p, c->part3.p);
        k_pack2<<<1, 1, 0, c->stream>>>(c->part2.p, c->part3.p);
        if (comm_allreduce_sum(c->comm, c->part2.p, 2, c->stream)) return set_err(comm_error());
        k_unpack2<<<1, 1, 0, c->stream>>>(c->part2.p, c->part3.p);
        V.n_part = 1;
        c->st.kernel_launches += 3;
    }
    k_cg_init_finish<<<1, VEC_THREADS, 0, c->stream>>>(V);
    c->st.kernel_launches++;
    CK(cudaGetLastError());
    const long long per_graph = c->graph_launches[method];
    // Chunks of CG_CHUNK captured iterations; the stop flag (and the peer-memory error flag) is copied to a pinned slot
    // after every chunk.  The next chunk is launched BEFORE the host looks at the previous one's flags, so the device never
    // waits for the host; once stopped, the kernels of a chunk return immediately.
    auto poll = [&](int slot) -> int {
        CK(cudaMemcpyAsync(&c->hPoll[slot].S, c->S.p, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
        if (c->comm.p2p)
            CK(cudaMemcpyAsync(&c->hPoll[slot].p2p_error, &((ArenaHdr*)c->comm.arena)->error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        else c->hPoll[slot].p2p_error = 0;
        CK(cudaEventRecord(c->evPoll[slot], c->stream));
        return 0;
    };
    int cur = 0;
    if (poll(cur)) return 1;
    CK(cudaEventSynchronize(c->evPoll[cur]));
    if (!c->hPoll[cur].S.stop) {
        CK(cudaGraphLaunch(c->graph[method], c->stream));
        c->st.kernel_launches += per_graph;
        cur ^= 1;
        if (poll(cur)) return 1;
        for (;;) {
            // one chunk (slot cur) is in flight: queue the next one, then look at the flags of the one in flight
            CK(cudaGraphLaunch(c->graph[method], c->stream));
            c->st.kernel_launches += per_graph;
            const int nxt = cur ^ 1;
            if (poll(nxt)) return 1;
            CK(cudaEventSynchronize(c->evPoll[cur]));
            const bool done = c->hPoll[cur].S.stop || c->hPoll[cur].p2p_error || c->hPoll[cur].S.iter >= max_cg_iterations;
            cur = nxt;
            if (done) break;
        }
        CK(cudaEventSynchronize(c->evPoll[cur]));
    }
    *c->hS = c->hPoll[cur].S;
    if (c->hPoll[cur].p2p_error) return set_err("peer-memory collective timed out (a rank stopped responding); the communicator must be rebuilt");
    return 0;
}

int en234gpu_solve_dev(en234gpu_handle c, int method, double cg_tolerance, int max_cg_iterations,
                       double* correction_norm, int* cg_iterations, int* fail) {
    if (method == EN234_SOLVE_BICGSTAB_BSR) return bicgstab_solve(c, cg_tolerance, max_cg_iterations, correction_norm, cg_iterations, fail);
    if (method != EN234_SOLVE_PCG_BSR && method != EN234_SOLVE_PCG_MATRIXFREE) return set_err("unknown solve method");
    if (method == EN234_SOLVE_PCG_BSR && !c->matrix_valid) return set_err("solve called before assemble");
    if (method != c->mode) return set_err("solve method differs from the operator mode (en234gpu_set_operator_mode)");
    if (c->any_c3d8) return set_err("C3D8 continuum elements have an unsymmetric tangent: use EN234_SOLVE_BICGSTAB_BSR");
    if (method == EN234_SOLVE_PCG_MATRIXFREE && c->any_neo)
        return set_err("the matrix-free operator covers the linear-elastic elements only (1001 / U1)");
    if (phase_begin(c)) return 1;
    if (c->ties.n && ties_before_solve(c)) return 1;
    if (pcg_run(c, method, cg_tolerance, max_cg_iterations)) return 1;
    *fail = c->hS->fail ? 1 : 0;
    *cg_iterations = c->hS->iter;
    c->st.last_cg_iterations = c->hS->iter;
    c->st.last_cg_err = c->hS->err;
    const int cg_iter = c->hS->iter;
    const double cg_err = c->hS->err;
    if (!*fail) {
        double dl2 = 0.0;
        if (c->ties.n && ties_after_solve(c, c->sol.p, &dl2)) return 1;
        k_add_solution<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(c->ndof, c->sol.p, c->du.p,
                                                                  c->have_owned ? c->owned.p : nullptr, c->part2.p);
        c->st.kernel_launches++;
        if (finish_norm(c, c->part2.p, c->grid_vec, correction_norm)) return 1;
        if (c->ties.n) *correction_norm = std::sqrt(*correction_norm * *correction_norm + dl2);   // multiplier corrections count (:971)
    } else *correction_norm = 0.0;
    c->hS->iter = cg_iter; c->hS->err = cg_err;
    if (phase_end(c, &c->st.solve_ms)) return 1;
    c->st.spmv_ms_avg = c->hS->iter > 0 ? c->st.solve_ms / c->hS->iter : 0.0;
    return 0;
}

// ---- state movement -------------------------------------------------------------------------------------
int en234gpu_upload_state(en234gpu_handle c, const double* dof_total, const double* dof_increment,
                          const double* f_element_ext, const double* f_ext, int n_fixed, const int* fixed_dof,
                          const double* fixed_value) {
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->ndof * sizeof(double);
    if (dof_total) CK(cudaMemcpyAsync(c->ut.p, dof_total, b, cudaMemcpyHostToDevice, c->stream));
    if (dof_increment) CK(cudaMemcpyAsync(c->du.p, dof_increment, b, cudaMemcpyHostToDevice, c->stream));
    c->have_felem = f_element_ext != nullptr;
    c->have_fext = f_ext != nullptr;
    if (f_element_ext) CK(cudaMemcpyAsync(c->felem.p, f_element_ext, b, cudaMemcpyHostToDevice, c->stream));
    if (f_ext) CK(cudaMemcpyAsync(c->fext.p, f_ext, b, cudaMemcpyHostToDevice, c->stream));
    if (n_fixed >= 0) return upload_fixed(c, n_fixed, fixed_dof, fixed_value);
    return 0;
}

int en234gpu_download_state(en234gpu_handle c, double* dof_increment, double* rforce) {
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->ndof * sizeof(double);
    if (dof_increment) CK(cudaMemcpyAsync(dof_increment, c->du.p, b, cudaMemcpyDeviceToHost, c->stream));
    if (rforce) CK(cudaMemcpyAsync(rforce, c->rforce.p, b, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

long long en234gpu_n_state_variables(en234gpu_handle c) { return (long long)c->sv0.n; }

int en234gpu_commit_state(en234gpu_handle c) {
    CK(cudaSetDevice(c->device));
    if (c->sv0.n) CK(cudaMemcpyAsync(c->sv0.p, c->sv1.p, c->sv0.n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int en234gpu_get_state_variables(en234gpu_handle c, int updated, double* out) {
    CK(cudaSetDevice(c->device));
    if (c->sv0.n) {
        CK(cudaMemcpyAsync(out, updated ? c->sv1.p : c->sv0.p, c->sv0.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int en234gpu_set_state_variables(en234gpu_handle c, const double* initial) {
    CK(cudaSetDevice(c->device));
    if (c->sv0.n) {
        CK(cudaMemcpyAsync(c->sv0.p, initial, c->sv0.n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int en234gpu_zero_increment_dev(en234gpu_handle c) {
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->du.p, 0, (size_t)c->ndof * sizeof(double), c->stream));
    return 0;
}

// ---- host-buffer entry points (what a Fortran driver binds) -------------------------------------------------
int en234gpu_assemble(en234gpu_handle c, const double* dof_total, const double* dof_increment,
                      const double* f_element_ext, double* rforce, double* nodal_force_norm, int* fail) {
    if (!dof_total || !dof_increment) return set_err("en234gpu_assemble: null DOF arrays");
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->ndof * sizeof(double);
    CK(cudaMemcpyAsync(c->ut.p, dof_total, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->du.p, dof_increment, b, cudaMemcpyHostToDevice, c->stream));
    c->have_felem = f_element_ext != nullptr;
    if (f_element_ext) CK(cudaMemcpyAsync(c->felem.p, f_element_ext, b, cudaMemcpyHostToDevice, c->stream));
    if (en234gpu_assemble_dev(c, nodal_force_norm, fail)) return 1;
    if (rforce) {
        CK(cudaMemcpyAsync(rforce, c->rforce.p, b, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int en234gpu_apply_boundaryconditions(en234gpu_handle c, const double* f_ext, int n_fixed, const int* fixed_dof,
                                      const double* fixed_correction, double* unbalanced_force_norm) {
    CK(cudaSetDevice(c->device));
    c->have_fext = f_ext != nullptr;
    if (f_ext) CK(cudaMemcpyAsync(c->fext.p, f_ext, (size_t)c->ndof * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (upload_fixed(c, n_fixed, fixed_dof, fixed_correction)) return 1;
    if (phase_begin(c)) return 1;
    if (apply_bc_common(c, 1, unbalanced_force_norm)) return 1;
    return phase_end(c, &c->st.bc_ms);
}

int en234gpu_solve(en234gpu_handle c, int method, double cg_tolerance, int max_cg_iterations, double* dof_increment,
                   double* correction_norm, int* cg_iterations, int* fail) {
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->ndof * sizeof(double);
    CK(cudaMemcpyAsync(c->du.p, dof_increment, b, cudaMemcpyHostToDevice, c->stream));
    if (en234gpu_solve_dev(c, method, cg_tolerance, max_cg_iterations, correction_norm, cg_iterations, fail)) return 1;
    CK(cudaMemcpyAsync(dof_increment, c->du.p, b, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- element plugin surface / inspection --------------------------------------------------------------------
int en234gpu_element_batch(en234gpu_handle c, int n, const int* element_numbers, const double* dof_total,
                           const double* dof_increment, double* element_stiffness, double* element_residual, int* fail) {
    if (c->npe != 8) return set_err("en234gpu_element_batch: hex8 meshes only");
    CK(cudaSetDevice(c->device));
    std::vector<int> el(n);
    for (int i = 0; i < n; ++i) {
        el[i] = element_numbers[i] - 1;
        if (el[i] < 0 || el[i] >= c->E) return set_err("element number out of range");
    }
    DevBuf<int> d_el;
    DevBuf<double> d_K, d_R;
    CK(d_el.alloc(n)); CK(d_K.alloc((size_t)n * 576)); CK(d_R.alloc((size_t)n * 24));
    const size_t b = (size_t)c->ndof * sizeof(double);
    CK(cudaMemcpyAsync(d_el.p, el.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tmpvec.p, dof_total, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tmpvec2.p, dof_increment, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(&c->S.p->nanflag, 0, sizeof(int), c->stream));
    k_element_batch<<<n, 64, 0, c->stream>>>(mesh_of(c), n, d_el.p, c->tmpvec.p, c->tmpvec2.p, d_K.p, d_R.p, c->S.p);
    c->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(element_stiffness, d_K.p, (size_t)n * 576 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(element_residual, d_R.p, (size_t)n * 24 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (read_scalars(c)) return 1;
    if (fail) *fail = c->hS->nanflag ? 1 : 0;
    return 0;
}

int en234gpu_get_svars(en234gpu_handle c, const double* dof_total, const double* dof_increment, double* svars) {
    if (c->npe != 8) return set_err("en234gpu_get_svars: hex8 meshes only");
    CK(cudaSetDevice(c->device));
    DevBuf<double> d_sv;
    CK(d_sv.alloc((size_t)c->E * 48));
    const size_t b = (size_t)c->ndof * sizeof(double);
    CK(cudaMemcpyAsync(c->tmpvec.p, dof_total, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tmpvec2.p, dof_increment, b, cudaMemcpyHostToDevice, c->stream));
    const long long nt = (long long)c->E * 8;
    k_svars<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(mesh_of(c), c->tmpvec.p, c->tmpvec2.p, d_sv.p);
    c->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(svars, d_sv.p, (size_t)c->E * 48 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// sum of a host buffer over the ranks of the handle's communicator (post-processing: gathers rank-local nodal fields into
// whole-mesh arrays for the Tecplot writer of a partitioned run); a handle without a communicator leaves the buffer unchanged
int en234gpu_comm_allreduce_host(en234gpu_handle c, double* buf, long long n) {
    if (!c->comm.active || n <= 0) return 0;
    if (!c->comm.nccl) return set_err("en234gpu_comm_allreduce_host: needs the NCCL communicator (en234gpu_comm_init)");
    CK(cudaSetDevice(c->device));
    const long long chunk = std::min<long long>(n, 1LL << 24);
    DevBuf<double> d;
    CK(d.alloc((size_t)chunk));
    for (long long o = 0; o < n; o += chunk) {
        const long long m = std::min(chunk, n - o);
        CK(cudaMemcpyAsync(d.p, buf + o, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (comm_allreduce_sum(c->comm, d.p, (int)m, c->stream)) return set_err(comm_error());
        CK(cudaMemcpyAsync(buf + o, d.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// print_state's projection of integration-point fields to the nodes (src/printing.f90:457-490): right-hand sides, the
// 27-point mass matrix written as diag(m,m,m) blocks over the stiffness storage, three fields per Jacobi-PCG pass (or the
// row-sum lumped matrix).  The stiffness values are overwritten: the next solve needs a fresh assembly.
int en234gpu_project_fields(en234gpu_handle c, int n_fields, const int* field_codes, int lumped, double tolerance,
                            int max_iterations, const double* dof_total, const double* dof_increment, double* fields,
                            int* iterations) {
    if (c->npe != 8) return set_err("en234gpu_project_fields: hex8 meshes only");
    CK(cudaSetDevice(c->device));
    // post-processing, not a timed path: every stage is synchronised so that a failure names the stage
#define STAGE(name)                                                                                              \
    do {                                                                                                         \
        cudaError_t e_ = cudaGetLastError();                                                                     \
        if (e_ == cudaSuccess) e_ = cudaStreamSynchronize(c->stream);                                            \
        if (e_ != cudaSuccess) return set_err(std::string("project_fields, ") + name + ": " + cudaGetErrorString(e_)); \
    } while (0)
    if (iterations) *iterations = 0;
    if (n_fields < 0 || n_fields > PROJ_MAXF) return set_err("project_fields: between 0 and " + std::to_string(PROJ_MAXF) + " field variables");
    if (c->any_neo)   // user_element_fieldvariables knows element 1001 only (user_codes/src/user_element.f90:233-266)
        return set_err("project_fields: no field variables are defined for the neo-Hookean element (1002 / U2)");
    if (n_fields == 0) return 0;
    const int N = c->N;
    const size_t b = (size_t)c->ndof * sizeof(double);
    if (dof_total) CK(cudaMemcpyAsync(c->ut.p, dof_total, b, cudaMemcpyHostToDevice, c->stream));
    if (dof_increment) CK(cudaMemcpyAsync(c->du.p, dof_increment, b, cudaMemcpyHostToDevice, c->stream));
    // partitioned runs: right-hand sides, matrix diagonal and lumped entries of interface nodes are partial sums over the
    // rank's elements; they are halo-summed as DOF-shaped vectors, the PCG loop is the partitioned one of the static step
    const bool part = c->comm.active;
    const int ls = part ? 3 : 1;
    DevBuf<double> F, lump;
    CK(F.alloc((size_t)n_fields * N));
    CK(lump.alloc((size_t)ls * N));
    ProjArgs A{};
    A.M = mesh_of(c);
    A.ut = c->ut.p; A.du = c->du.p;
    A.source = c->any_c3d8 ? 1 : 0;
    A.sv = c->any_c3d8 ? c->sv1.p : nullptr;
    A.nsp = C3D8_NSV_POINT;
    A.nf = n_fields;
    A.F = F.p;
    for (int k = 0; k < n_fields; ++k) {
        int code = field_codes[k];
        bool ok;
        if (A.source == 0) ok = code >= 0 && code <= FIELD_MISES;                       // el_linelast_3Dbasic.f90:392-408
        else ok = (code >= 0 && code < 6) || (code >= FIELD_STRAIN0 && code < FIELD_STRAIN0 + 6) ||
                  (code >= FIELD_STATE0 && code < FIELD_STATE0 + A.nsp);                // continuum_elements.f90:1291-1345
        A.code[k] = ok ? code : FIELD_ZERO;                                             // unknown names stay zero fields
    }
    const int grid = (N + PROJ_THREADS - 1) / PROJ_THREADS;
    STAGE("state upload");
    k_project_rhs<<<grid, PROJ_THREADS, 0, c->stream>>>(A);
    c->st.kernel_launches++;
    STAGE("k_project_rhs");
    // the PCG graph is wired to rhs / minv / val: keep the step's rhs and preconditioner aside
    CK(cudaMemcpyAsync(c->tmpvec.p, c->rhs.p, b, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tmpvec2.p, c->minv.p, b, cudaMemcpyDeviceToDevice, c->stream));
    MassArgs Ma{};
    Ma.M = mesh_of(c); Ma.val = c->val.p; Ma.minv = c->minv.p; Ma.lump = lump.p;
    Ma.raw_diag = part ? 1 : 0; Ma.lump_stride = ls;
    c->matrix_valid = false;
    c->diag_valid = false;
    STAGE("save rhs / preconditioner");
    k_project_mass<<<grid, PROJ_THREADS, 0, c->stream>>>(Ma);
    c->st.kernel_launches++;
    STAGE("k_project_mass");
    if (part) {
        if (comm_halo_sum(c->comm, c->minv.p, c->stream) || comm_halo_sum(c->comm, lump.p, c->stream)) return set_err(comm_error());
        k_project_invert<<<(c->ndof + 255) / 256, 256, 0, c->stream>>>(c->ndof, c->minv.p);
        c->st.kernel_launches += 1 + 2 * c->comm.launches_per_halo;
        STAGE("halo sums of the projection matrix diagonal");
    }
    int rc = 0, its = 0;
    if (!lumped || part) {
        for (int k0 = 0; k0 < n_fields && !rc; k0 += 3) {
            k_project_pack<<<(N + 255) / 256, 256, 0, c->stream>>>(N, n_fields, k0, F.p, c->rhs.p);
            c->st.kernel_launches++;
            if (part && comm_halo_sum(c->comm, c->rhs.p, c->stream)) { rc = set_err(comm_error()); break; }
            STAGE("k_project_pack");
            if (lumped) {                     // lumped: only the (halo-summed) right-hand sides are needed
                if (part) {
                    k_project_unpack<<<(N + 255) / 256, 256, 0, c->stream>>>(N, n_fields, k0, c->rhs.p, F.p);
                    c->st.kernel_launches++;
                }
                continue;
            }
            rc = pcg_run(c, EN234_SOLVE_PCG_BSR, tolerance > 0 ? tolerance : 1e-28, max_iterations > 0 ? max_iterations : 500);
            if (rc) break;
            if (c->hS->fail) { rc = set_err("project_fields: the projection solve did not converge"); break; }
            its += c->hS->iter;
            k_project_unpack<<<(N + 255) / 256, 256, 0, c->stream>>>(N, n_fields, k0, c->sol.p, F.p);
            c->st.kernel_launches++;
            STAGE("k_project_unpack");
        }
    }
    if (lumped && !rc) {
        k_project_lumped<<<(N + 255) / 256, 256, 0, c->stream>>>(N, n_fields, lump.p, ls, F.p);
        c->st.kernel_launches++;
        STAGE("k_project_lumped");
    }
    CK(cudaMemcpyAsync(c->rhs.p, c->tmpvec.p, b, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->minv.p, c->tmpvec2.p, b, cudaMemcpyDeviceToDevice, c->stream));
    if (rc) { cudaStreamSynchronize(c->stream); return 1; }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(fields, F.p, (size_t)n_fields * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (iterations) *iterations = its;
    return 0;
#undef STAGE
}

// ---- explicit dynamic step (src/explicit_dynamic_step.f90) ---------------------------------------------------
constexpr int DYN_CHUNK = 16;      // time steps per CUDA-graph launch

int en234gpu_dynamic_setup(en234gpu_handle c, const double* element_density, int n_histories, const int* history_ptr,
                           const double* history_time, const double* history_value, int n_loads, const int* load_ptr,
                           const int* load_dof, const double* load_val, const int* load_history, const double* load_constant,
                           int n_pdof, const int* pdof_dof, const int* pdof_history, const double* pdof_value,
                           const int* pdof_mode) {
    if (c->npe != 8) return set_err("explicit dynamics: hex8 meshes only");
    CK(cudaSetDevice(c->device));
    if (c->any_neo || c->any_c3d8)   // user_element_dynamic knows element 1001 only (user_element.f90:152-181); VUMAT is out of scope
        return set_err("explicit dynamics covers the linear elastic hex8 elements (1001 / VU1) only");
    auto& D = c->dyn;
    const size_t nd = (size_t)c->ndof;
#define UPD(buf, ptr, count)                                                                          \
    do {                                                                                              \
        CK(D.buf.alloc(std::max<size_t>((count), 1)));                                                \
        if ((count) > 0) CK(cudaMemcpy(D.buf.p, (ptr), (size_t)(count) * sizeof(*(ptr)), cudaMemcpyHostToDevice)); \
    } while (0)
    UPD(density, element_density, (size_t)c->E);
    const int nh = std::max(n_histories, 0);
    std::vector<int> hp(1, 0);
    if (nh > 0) hp.assign(history_ptr, history_ptr + nh + 1);
    UPD(hist_ptr, hp.data(), hp.size());
    UPD(hist_t, history_time, (size_t)hp.back());
    UPD(hist_v, history_value, (size_t)hp.back());
    D.load_ptr.assign(1, 0); D.load_hist.clear(); D.load_const.clear();
    if (n_loads > 0) {
        D.load_ptr.assign(load_ptr, load_ptr + n_loads + 1);
        D.load_hist.assign(load_history, load_history + n_loads);
        D.load_const.assign(load_constant, load_constant + n_loads);
        for (int l = 0; l < n_loads; ++l)
            if (D.load_hist[l] < 0 || D.load_hist[l] > nh) return set_err("dynamic_setup: load refers to a history that does not exist");
        for (int k = 0; k < D.load_ptr.back(); ++k)
            if (load_dof[k] < 0 || (size_t)load_dof[k] >= nd) return set_err("dynamic_setup: load DOF out of range");
    }
    UPD(load_dof, load_dof, (size_t)D.load_ptr.back());
    UPD(load_val, load_val, (size_t)D.load_ptr.back());
    for (int k = 0; k < n_pdof; ++k) {
        if (pdof_dof[k] < 0 || (size_t)pdof_dof[k] >= nd) return set_err("dynamic_setup: prescribed DOF out of range");
        if (pdof_history[k] < 0 || pdof_history[k] > nh) return set_err("dynamic_setup: prescribed DOF refers to a history that does not exist");
    }
    D.n_pd = n_pdof;
    UPD(pd_dof, pdof_dof, (size_t)n_pdof); UPD(pd_hist, pdof_history, (size_t)n_pdof);
    UPD(pd_value, pdof_value, (size_t)n_pdof); UPD(pd_mode, pdof_mode, (size_t)n_pdof);
#undef UPD
    for (DevBuf<double>* b : {&D.v, &D.a, &D.mass, &D.f}) {
        CK(b->alloc(nd));
        CK(cudaMemsetAsync(b->p, 0, nd * sizeof(double), c->stream));
    }
    CK(D.time.alloc(1));
    CK(cudaMemsetAsync(D.time.p, 0, sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->ut.p, 0, nd * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->du.p, 0, nd * sizeof(double), c->stream));
    k_dyn_lumped_mass<<<(c->N + 127) / 128, 128, 0, c->stream>>>(mesh_of(c), D.density.p, D.mass.p);
    c->st.kernel_launches++;
    CK(cudaGetLastError());
    // partitioned runs: interface nodes hold the mass of the local elements only
    if (c->comm.active && comm_halo_sum(c->comm, D.mass.p, c->stream)) return set_err(comm_error());
    CK(cudaStreamSynchronize(c->stream));
    if (c->Re.n < (size_t)c->E * 24) CK(c->Re.alloc((size_t)c->E * 24));
    mf_prepare_geom(c->mf, c->E);
    if (D.graph) { cudaGraphExecDestroy(D.graph); D.graph = nullptr; }
    D.ready = true;
    c->matrix_valid = false;
    c->rhs_valid = false;
    return 0;
}

int en234gpu_dynamic_set_state(en234gpu_handle c, const double* dof_total, const double* velocity, const double* acceleration,
                               double time) {
    CK(cudaSetDevice(c->device));
    if (!c->dyn.ready) return set_err("dynamic_set_state called before dynamic_setup");
    const size_t b = (size_t)c->ndof * sizeof(double);
    auto put = [&](double* dst, const double* src) -> cudaError_t {
        return src ? cudaMemcpyAsync(dst, src, b, cudaMemcpyHostToDevice, c->stream) : cudaMemsetAsync(dst, 0, b, c->stream);
    };
    CK(put(c->ut.p, dof_total)); CK(put(c->dyn.v.p, velocity)); CK(put(c->dyn.a.p, acceleration));
    CK(cudaMemcpyAsync(c->dyn.time.p, &time, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// phases: 1 = predictor + boundary conditions, 2 = element forces + update, 3 = both (one whole step)
static int dyn_enqueue_step(en234gpu_ctx* c, double dt, int phases = 3) {
    auto& D = c->dyn;
    const int n = c->ndof;
    const DynTables T{D.hist_ptr.p, D.hist_t.p, D.hist_v.p};
    if (phases & 1) {
    k_dyn_predict<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(n, dt, D.v.p, D.a.p, c->du.p, D.f.p);
    const int n_loads = (int)D.load_hist.size();
    for (int l = 0; l < n_loads; ++l) {
        const int cnt = D.load_ptr[l + 1] - D.load_ptr[l];
        if (cnt == 0) continue;
        k_dyn_load<<<(cnt + 255) / 256, 256, 0, c->stream>>>(cnt, D.load_dof.p + D.load_ptr[l], D.load_val.p + D.load_ptr[l],
                                                             D.load_hist[l], D.load_const[l], T, D.time.p, dt, D.f.p);
        c->st.kernel_launches++;
    }
    if (D.n_pd > 0) {
        k_dyn_pdof<<<(D.n_pd + 255) / 256, 256, 0, c->stream>>>(D.n_pd, D.pd_dof.p, D.pd_hist.p, D.pd_value.p, D.pd_mode.p, T,
                                                                D.time.p, dt, c->ut.p, c->du.p);
        c->st.kernel_launches++;
    }
    c->st.kernel_launches++;
    }
    if (!(phases & 2)) return 0;
    ElemArgs EA;
    EA.M = mesh_of(c);
    EA.Ke = nullptr; EA.Re = c->Re.p; EA.S = c->S.p;
    launch_element_forces(c->mf, EA, LoadSum{c->ut.p, c->du.p}, c->grid_elem0, c->stream);
    if (c->comm.active) {
        // partial element-force sums -> halo sum -> update (the one exchange of a step)
        GatherArgs G{};
        G.M = EA.M; G.Re = c->Re.p; G.rhs = c->rhs.p; G.rforce = c->rforce.p; G.partials = c->part.p; G.S = c->S.p;
        const int np = std::min(MAX_PARTIALS, std::max(1, (c->N + VEC_THREADS - 1) / VEC_THREADS));
        k_gather_residual<<<np, VEC_THREADS, 0, c->stream>>>(G);
        if (comm_halo_sum(c->comm, c->rhs.p, c->stream)) return 1;
        k_dyn_update<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(n, dt, D.f.p, c->rhs.p, D.mass.p, D.a.p, D.v.p, c->ut.p, c->du.p,
                                                                 c->rforce.p, D.time.p, c->S.p);
        c->st.kernel_launches += 3 + c->comm.launches_per_halo;
        return 0;
    }
    DynUpdateArgs U;
    U.M = EA.M; U.Re = c->Re.p; U.f = D.f.p; U.mass = D.mass.p; U.a = D.a.p; U.v = D.v.p; U.ut = c->ut.p; U.rforce = c->rforce.p;
    U.du = c->du.p; U.time = D.time.p; U.dt = dt; U.S = c->S.p;
    k_dyn_gather_update<<<c->grid_vec, VEC_THREADS, 0, c->stream>>>(U);
    c->st.kernel_launches += 2;
    return 0;
}

// One step in two halves, so that the caller can read (dof_total, dof_increment) where the reference prints the state:
// after the boundary conditions, before the element forces (explicit_dynamic_step.f90:40-64).  phase 1, then phase 2.
int en234gpu_dynamic_half_step(en234gpu_handle c, double dt, int phase, int* fail) {
    CK(cudaSetDevice(c->device));
    if (!c->dyn.ready) return set_err("dynamic_half_step called before dynamic_setup");
    if (phase != 1 && phase != 2) return set_err("dynamic_half_step: phase is 1 (predictor + boundary conditions) or 2 (forces + update)");
    if (!(dt > 0.0)) return set_err("dynamic_half_step: dt must be positive");
    if (phase == 1) CK(cudaMemsetAsync(&c->S.p->nanflag, 0, sizeof(int), c->stream));
    if (dyn_enqueue_step(c, dt, phase)) return set_err(comm_error());
    CK(cudaGetLastError());
    if (read_scalars(c)) return 1;
    int bad = c->hS->nanflag;
    if (c->comm.active && comm_host_max(c->comm, &bad, c->stream)) return set_err(comm_error());
    if (fail) *fail = bad ? 1 : 0;
    return 0;
}

// n_steps central-difference steps of size dt from the device-resident state; fail = 1 when a non-finite acceleration or
// a non-positive Jacobian appeared.  device_ms (may be NULL): CUDA-event time of the steps.
int en234gpu_dynamic_steps(en234gpu_handle c, double dt, int n_steps, int* fail, double* device_ms) {
    CK(cudaSetDevice(c->device));
    auto& D = c->dyn;
    if (!D.ready) return set_err("dynamic_steps called before dynamic_setup");
    if (!(dt > 0.0) || n_steps < 0) return set_err("dynamic_steps: dt must be positive");
    CK(cudaMemsetAsync(&c->S.p->nanflag, 0, sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    int left = n_steps;
    if (left >= 2 * DYN_CHUNK) {
        if (D.graph && D.graph_dt != dt) { cudaGraphExecDestroy(D.graph); D.graph = nullptr; }
        if (!D.graph) {
            cudaGraph_t g = nullptr;
            const long long before = c->st.kernel_launches;
            CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int rc = 0;
            for (int s = 0; s < DYN_CHUNK && !rc; ++s) rc = dyn_enqueue_step(c, dt);
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            D.graph_launches = c->st.kernel_launches - before;
            c->st.kernel_launches = before;
            if (rc) return set_err(std::string("dynamic_steps: graph capture: ") + comm_error());
            if (e != cudaSuccess) return set_err(std::string("dynamic_steps: cudaStreamEndCapture: ") + cudaGetErrorString(e));
            CK(cudaGraphInstantiate(&D.graph, g, 0));
            CK(cudaGraphDestroy(g));
            D.graph_dt = dt;
        }
        for (; left >= DYN_CHUNK; left -= DYN_CHUNK) {
            CK(cudaGraphLaunch(D.graph, c->stream));
            c->st.kernel_launches += D.graph_launches;
        }
    }
    for (; left > 0; --left)
        if (dyn_enqueue_step(c, dt)) return set_err(comm_error());
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));
    if (read_scalars(c)) return 1;
    int bad = c->hS->nanflag;
    if (c->comm.active && comm_host_max(c->comm, &bad, c->stream)) return set_err(comm_error());
    if (fail) *fail = bad ? 1 : 0;
    if (device_ms) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *device_ms = ms;
    }
    return 0;
}

int en234gpu_dynamic_get_state(en234gpu_handle c, double* dof_total, double* dof_increment, double* velocity,
                               double* acceleration, double* rforce, double* lumped_mass, double* time) {
    CK(cudaSetDevice(c->device));
    auto& D = c->dyn;
    if (!D.ready) return set_err("dynamic_get_state called before dynamic_setup");
    const size_t b = (size_t)c->ndof * sizeof(double);
    auto get = [&](double* dst, const double* src) -> cudaError_t {
        return dst ? cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
    };
    CK(get(dof_total, c->ut.p)); CK(get(dof_increment, c->du.p)); CK(get(velocity, D.v.p)); CK(get(acceleration, D.a.p));
    CK(get(rforce, c->rforce.p)); CK(get(lumped_mass, D.mass.p));
    if (time) CK(cudaMemcpyAsync(time, D.time.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// Single-element evaluation on the device for the host-callable twins of the reference's element routines
// (user_element_static / UEL): an 8-node scratch mesh is kept on device 0 and refilled per call.
int en234gpu_element_single(int element_flag, const double* coords24, const double* props, int n_props,
                            const double* dof_total24, const double* dof_increment24, double* element_stiffness,
                            double* element_residual, double* svars48, double* energy, int* fail) {
    static en234gpu_handle scratch = nullptr;
    static int s_flag = 0;
    static double s_props[2] = {0, 0}, s_coords[24];
    bool fresh = scratch == nullptr || s_flag != element_flag || n_props < 2 || s_props[0] != props[0] ||
                 s_props[1] != props[1] || std::memcmp(s_coords, coords24, sizeof(s_coords)) != 0;
    if (fresh) {
        if (scratch) { en234gpu_destroy(scratch); scratch = nullptr; }
        const int conn[8] = {1, 2, 3, 4, 5, 6, 7, 8}, pidx = 0;
        if (en234gpu_create(&scratch, 0, 8, 1, coords24, conn, &element_flag, &pidx, props, n_props)) return 1;
        s_flag = element_flag; s_props[0] = props[0]; s_props[1] = props[1];
        std::memcpy(s_coords, coords24, sizeof(s_coords));
    }
    en234gpu_ctx* c = scratch;
    CK(cudaSetDevice(c->device));
    DevBuf<int> d_el;
    DevBuf<double> d_K, d_R, d_sv, d_en;
    const int zero = 0;
    CK(d_el.alloc(1)); CK(d_K.alloc(576)); CK(d_R.alloc(24)); CK(d_sv.alloc(48)); CK(d_en.alloc(1));
    CK(cudaMemcpyAsync(d_el.p, &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tmpvec.p, dof_total24, 24 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tmpvec2.p, dof_increment24, 24 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(&c->S.p->nanflag, 0, sizeof(int), c->stream));
    k_element_batch<<<1, 64, 0, c->stream>>>(mesh_of(c), 1, d_el.p, c->tmpvec.p, c->tmpvec2.p, d_K.p, d_R.p, c->S.p, d_sv.p, d_en.p);
    c->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(element_stiffness, d_K.p, 576 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(element_residual, d_R.p, 24 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (svars48) CK(cudaMemcpyAsync(svars48, d_sv.p, 48 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (energy) CK(cudaMemcpyAsync(energy, d_en.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (read_scalars(c)) return 1;
    if (fail) *fail = c->hS->nanflag ? 1 : 0;
    return 0;
}

long long en234gpu_nnz_blocks(en234gpu_handle c) { return c->nnzb; }

int en234gpu_get_csr(en234gpu_handle c, long long* rowptr, int* col, double* val) {
    CK(cudaSetDevice(c->device));
    // block row i stores [r][block][c]: scalar row 3i+r starts at 9*browptr[i] + r*3*nb
    for (int i = 0; i < c->N; ++i) {
        const long long rb = c->h_browptr[i], nb = c->h_browptr[i + 1] - rb;
        for (int r = 0; r < 3; ++r) {
            const long long start = 9 * rb + r * 3 * nb;
            if (rowptr) rowptr[3 * (long long)i + r] = start;
            if (col)
                for (long long q = 0; q < 3 * nb; ++q) col[start + q] = 3 * c->h_bcol[rb + q / 3] + (int)(q % 3);
        }
    }
    if (rowptr) rowptr[3 * (long long)c->N] = 9 * c->nnzb;
    if (val) {
        CK(cudaMemcpyAsync(val, c->val.p, 9 * (size_t)c->nnzb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int en234gpu_get_rhs(en234gpu_handle c, double* rhs, double* diag) {
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->ndof * sizeof(double);
    if (rhs) CK(cudaMemcpyAsync(rhs, c->rhs.p, b, cudaMemcpyDeviceToHost, c->stream));
    if (diag) {
        CK(cudaMemcpyAsync(diag, c->minv.p, b, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int d = 0; d < c->ndof; ++d) diag[d] = 1.0 / diag[d];
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int en234gpu_apply_operator(en234gpu_handle c, int method, const double* x, double* y) {
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->ndof * sizeof(double);
    CK(cudaMemcpyAsync(c->tmpvec.p, x, b, cudaMemcpyHostToDevice, c->stream));
    if (launch_operator(c, method, c->tmpvec.p, c->tmpvec2.p, 0)) return 1;
    CK(cudaGetLastError());
    if (c->comm.active && c->comm.p2p) { p2p_halo_fork(c, c->tmpvec2.p, 0); p2p_halo_join(c); }
    else if (c->comm.active && comm_halo_sum(c->comm, c->tmpvec2.p, c->stream)) return set_err(comm_error());
    CK(cudaMemcpyAsync(y, c->tmpvec2.p, b, cudaMemcpyDeviceToHost, c->stream));
    return read_scalars(c);        // synchronises and reports a timed-out halo exchange
}

// ---- multi-GPU ----------------------------------------------------------------------------------------------
int en234gpu_get_unique_id(char* id128) { return comm_unique_id(id128) ? set_err(comm_error()) : 0; }

int en234gpu_comm_init(en234gpu_handle c, int n_ranks, int rank, const char* id128) {
    if (c->npe != 8) return set_err("element partitions: hex8 meshes only");
    CK(cudaSetDevice(c->device));
    if (comm_init(c->comm, n_ranks, rank, id128, c->stream)) return set_err(comm_error());
    for (auto& g : c->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    return 0;
}

int en234gpu_set_partition(en234gpu_handle c, int n_neighbours, const int* neighbour_rank, const int* shared_ptr,
                           const int* shared_nodes, const unsigned char* owned) {
    CK(cudaSetDevice(c->device));
    CK(c->owned.alloc(c->N));
    CK(cudaMemcpy(c->owned.p, owned, c->N, cudaMemcpyHostToDevice));
    c->have_owned = true;
    if (comm_set_halo(c->comm, n_neighbours, neighbour_rank, shared_ptr, shared_nodes)) return set_err(comm_error());
    return 0;
}

int en234gpu_comm_arena(en234gpu_handle c, int max_shared_nodes, char* handle64) {
    CK(cudaSetDevice(c->device));
    const char* env = getenv("EN234_COMM");
    if (env && std::string(env) == "nccl") return set_err("peer-memory path disabled by EN234_COMM=nccl");
    if (comm_arena_create(c->comm, max_shared_nodes, handle64)) return set_err(comm_error());
    return 0;
}

int en234gpu_comm_connect(en234gpu_handle c, const char* handles) {
    CK(cudaSetDevice(c->device));
    if (comm_arena_connect(c->comm, handles)) return set_err(comm_error());
    for (auto& g : c->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    return 0;
}

int en234gpu_set_operator_mode(en234gpu_handle c, int method) {
    if (method != EN234_SOLVE_PCG_BSR && method != EN234_SOLVE_PCG_MATRIXFREE) return set_err("unknown operator mode");
    if (c->npe != 8 && method != EN234_SOLVE_PCG_BSR) return set_err("the matrix-free operators are hex8 kernels");
    c->mode = method;
    c->matrix_valid = false;
    c->rhs_valid = false;
    return 0;
}

int en234gpu_get_stats(en234gpu_handle c, struct en234gpu_stats* s) { c->st.mf_congruent = c->mf.congruent ? 1 : 0; *s = c->st; return 0; }

int en234gpu_measure_fp64_peak(en234gpu_handle c, double* tflops) {
    CK(cudaSetDevice(c->device));
    const int iters = 4096, grid = c->n_sm * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(c->ev0, c->stream));
        k_fp64_peak<<<grid, 256, 0, c->stream>>>(iters, 0.999999, c->tmpvec.p);
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaEventSynchronize(c->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->st.kernel_launches++;
        if (rep > 0) best = std::max(best, 2.0 * 16.0 * iters * (double)grid * 256.0 / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return 0;
}

// Latency of the collectives of the partitioned PCG loop in isolation (collective call: every rank, same arguments):
// which = 0: scalar all-reduce kernel (k_p2p_allreduce over grid_vec partials), 1: halo sum of a DOF vector (push + wait),
// 2: an empty single-thread kernel (launch floor).  Back-to-back launches inside one captured graph, like the loop.
int en234gpu_bench_collective(en234gpu_handle c, int which, int reps, double* us_per_call) {
    CK(cudaSetDevice(c->device));
    if (!c->comm.active || !c->comm.p2p) return set_err("bench_collective: needs a partitioned handle with connected peer-memory arenas");
    if (reps < 1) reps = 1;
    CK(cudaMemsetAsync(c->part2.p, 0, c->part2.n * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->part3.p, 0, c->part3.n * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->S.p, 0, sizeof(CgScalars), c->stream));
    CK(cudaMemsetAsync(c->Ap.p, 0, (size_t)c->ndof * sizeof(double), c->stream));
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 64; ++i) {
        if (which == 0) k_p2p_allreduce<<<1, VEC_THREADS, 0, c->stream>>>(c->comm.peers, c->part2.p, c->grid_vec, c->part3.p, c->grid_vec,
                                                                         c->part2.p, c->part3.p, c->S.p, 0);
        else if (which == 1) comm_halo_sum_p2p(c->comm, c->Ap.p, c->stream);
        else k_pack2<<<1, 1, 0, c->stream>>>(c->part2.p, c->part3.p);
    }
    CK(cudaStreamEndCapture(c->stream, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphDestroy(g));
    CK(cudaGraphLaunch(ge, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaGetLastError());
    float t = 0;
    CK(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    CK(cudaGraphExecDestroy(ge));
    *us_per_call = 1e3 * (double)t / (64.0 * reps);
    return 0;
}

int en234gpu_bench_operator(en234gpu_handle c, int method, int warm, int reps, double* ms_per_launch) {
    CK(cudaSetDevice(c->device));
    if (method == EN234_SOLVE_PCG_BSR && !c->matrix_valid) return set_err("bench_operator: no assembled matrix");
    for (int i = 0; i < warm; ++i) launch_operator(c, method, c->p.p, c->Ap.p, 0);
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) launch_operator(c, method, c->p.p, c->Ap.p, 0);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaGetLastError());
    float t = 0;
    CK(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    *ms_per_launch = (double)t / reps;
    return 0;
}

}  // extern "C"

#include "en234gpu_ties_host.h"
#include "en234gpu_multi.h"
