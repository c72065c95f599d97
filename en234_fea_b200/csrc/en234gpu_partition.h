// Element partition of a hex8 mesh for multi-GPU runs (host code, no CUDA).  The reference is serial;
// this is new.  Elements are split into n_parts contiguous ranges of the element numbering — whole layers when the mesh is
// numbered layer by layer (the structured blocks of user_mesh(): x fastest, then y, then z, so a range of layers is a slab).
// Each part keeps the closure of its elements' nodes: the nodes shared with other parts first ("interface" rows, in
// ascending global order), then the interior ones; a shared node is owned by the lowest part that holds it; the list of
// nodes shared with a neighbour is in ascending global node order on both sides, neighbours in ascending part number.
#pragma once
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

namespace en234part {

struct Part {
    long long e0 = 0, e1 = 0;              // element range [e0, e1)
    std::vector<int> l2g;                  // local node -> global node (0-based)
    int n_interface = 0;
    std::vector<int> conn;                 // 8 per local element, 1-based LOCAL node numbers
    std::vector<int> nbr, shared_ptr, shared_nodes;
    std::vector<unsigned char> owned;
};

// first element of every part (n_parts + 1 entries); empty string on success
inline std::string element_ranges(int n_nodes, long long E, const int* conn /*1-based, 8 per element*/, int n_parts,
                                  std::vector<long long>& first) {
    (void)n_nodes;
    if (n_parts < 1) return "partition: the number of parts must be positive";
    if (E < n_parts) return "partition: fewer elements (" + std::to_string(E) + ") than parts (" + std::to_string(n_parts) + ")";
    // layered numbering: the first element's upper face starts one node layer above its lower face; the elements of the
    // first layer are those whose first node lies inside the first node layer
    long long layer = 1;
    const int nodes_per_layer = conn[4] - conn[0];
    if (nodes_per_layer > 0) {
        long long cnt = 0;
        while (cnt < E && conn[8 * cnt] <= nodes_per_layer) ++cnt;
        if (cnt > 0 && E % cnt == 0 && E / cnt >= n_parts) layer = cnt;
    }
    const long long nl = E / layer;
    first.assign(n_parts + 1, 0);
    for (int r = 0; r <= n_parts; ++r) first[r] = nl * r / n_parts * layer;
    first[n_parts] = E;
    for (int r = 0; r < n_parts; ++r)
        if (first[r + 1] <= first[r]) return "partition: part " + std::to_string(r) + " would be empty";
    return "";
}

inline std::string partition(int n_nodes, long long E, const int* conn, int n_parts, std::vector<Part>& parts) {
    std::vector<long long> first;
    std::string err = element_ranges(n_nodes, E, conn, n_parts, first);
    if (!err.empty()) return err;
    for (long long q = 0; q < 8 * E; ++q)
        if (conn[q] < 1 || conn[q] > n_nodes) return "partition: connectivity refers to a node that does not exist";
    parts.assign(n_parts, Part());
    // (node, part) incidences of the nodes held by more than one part
    std::vector<int> rmin(n_nodes, n_parts), rmax(n_nodes, -1);
    for (int r = 0; r < n_parts; ++r)
        for (long long q = 8 * first[r]; q < 8 * first[r + 1]; ++q) {
            const int n = conn[q] - 1;
            rmin[n] = std::min(rmin[n], r);
            rmax[n] = std::max(rmax[n], r);
        }
    std::vector<std::pair<int, int>> inc;
    {
        std::vector<int> last(n_nodes, -1);
        for (int r = 0; r < n_parts; ++r)
            for (long long q = 8 * first[r]; q < 8 * first[r + 1]; ++q) {
                const int n = conn[q] - 1;
                if (rmin[n] != rmax[n] && last[n] != r) { last[n] = r; inc.emplace_back(n, r); }
            }
        std::sort(inc.begin(), inc.end());
    }
    std::vector<int> g2l(n_nodes), mark(n_nodes, -1);
    for (int r = 0; r < n_parts; ++r) {
        Part& P = parts[r];
        P.e0 = first[r]; P.e1 = first[r + 1];
        for (long long q = 8 * P.e0; q < 8 * P.e1; ++q) mark[conn[q] - 1] = r;
        for (int n = 0; n < n_nodes; ++n) if (mark[n] == r && rmin[n] != rmax[n]) { g2l[n] = (int)P.l2g.size(); P.l2g.push_back(n); }
        P.n_interface = (int)P.l2g.size();
        for (int n = 0; n < n_nodes; ++n) if (mark[n] == r && rmin[n] == rmax[n]) { g2l[n] = (int)P.l2g.size(); P.l2g.push_back(n); }
        P.conn.resize((size_t)8 * (P.e1 - P.e0));
        for (long long q = 8 * P.e0; q < 8 * P.e1; ++q) P.conn[q - 8 * P.e0] = g2l[conn[q] - 1] + 1;
        P.owned.resize(P.l2g.size());
        for (size_t i = 0; i < P.l2g.size(); ++i) P.owned[i] = rmin[P.l2g[i]] == r ? 1 : 0;
        // shared lists: for every other part q the interface nodes also held by q, in ascending global order
        std::vector<std::vector<int>> shared(n_parts);
        for (int i = 0; i < P.n_interface; ++i) {
            const int n = P.l2g[i];
            auto it = std::lower_bound(inc.begin(), inc.end(), std::make_pair(n, -1));
            for (; it != inc.end() && it->first == n; ++it)
                if (it->second != r) shared[it->second].push_back(i);
        }
        P.shared_ptr.assign(1, 0);
        for (int q = 0; q < n_parts; ++q)
            if (!shared[q].empty()) {
                P.nbr.push_back(q);
                P.shared_nodes.insert(P.shared_nodes.end(), shared[q].begin(), shared[q].end());
                P.shared_ptr.push_back((int)P.shared_nodes.size());
            }
    }
    return "";
}

}  // namespace en234part
