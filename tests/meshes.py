"""Test meshes of the other 3-D shapes of element 1001 — 4-node / 10-node tetrahedra, 20-node bricks — as literal `.in` text
(so the parser path is exercised): a jittered block whose face x = 0 is clamped and whose face x = max carries nodal forces.
Node orders are the reference's (user_codes/modules/Element_Utilities.f90:863-1025): tet10 mid-edge nodes 5..10 on edges
1-2, 2-3, 3-1, 1-4, 2-4, 3-4; hex20 mid-edge nodes 9..20 on 1-2, 2-3, 3-4, 4-1, 5-6, 6-7, 7-8, 8-5, 1-5, 2-6, 3-7, 4-8."""
import numpy as np

KUHN = [(0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)]   # corners as x + 2 y + 4 z bits
TET_EDGES = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
HEX_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def block(shape, nx, ny, nz, jitter=0.15, seed=7):
    """(coords (N, 3), connectivity (E, nodes_per_element) 1-based, nodes on x = 0, nodes on x = max)"""
    rng = np.random.default_rng(seed)
    g = np.zeros((nz + 1, ny + 1, nx + 1, 3))
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                p = np.array([i, j, k], float)
                if 0 < i < nx and 0 < j < ny and 0 < k < nz:
                    p += jitter * rng.uniform(-1, 1, 3)
                g[k, j, i] = p
    xyz = [tuple(p) for p in g.reshape(-1, 3)]
    nid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    mids = {}

    def mid(a, b):
        key = (min(a, b), max(a, b))
        if key not in mids:
            mids[key] = len(xyz)
            xyz.append(tuple(0.5 * (np.array(xyz[a]) + np.array(xyz[b]))))
        return mids[key]

    conn = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                if shape == "hex20":
                    c = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                         nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)]
                    conn.append(c + [mid(c[a], c[b]) for a, b in HEX_EDGES])
                    continue
                corner = lambda q: nid(i + (q & 1), j + ((q >> 1) & 1), k + ((q >> 2) & 1))
                for t in KUHN:
                    c = [corner(q) for q in t]
                    X = np.array([xyz[n] for n in c])
                    if np.linalg.det((X[:3] - X[3]).T) < 0:           # the reference's tet: dx/dxi_j = x_j - x_4
                        c[0], c[1] = c[1], c[0]
                    if shape == "tet10":
                        c = c + [mid(c[a], c[b]) for a, b in TET_EDGES]
                    conn.append(c)
    X = np.array(xyz)
    xmin = [n + 1 for n in range(len(xyz)) if abs(X[n, 0]) < 1e-12]
    xmax = [n + 1 for n in range(len(xyz)) if abs(X[n, 0] - nx) < 1e-12]
    return X, np.array(conn) + 1, xmin, xmax


def input_text(shape, nx, ny, nz, jitter=0.15, E=100.0, nu=0.3, solver="CONJUGATE GRADIENT", cg_tolerance=1e-26, cg_maxit=20000,
               force=(0.0, 0.0, -0.01)):
    X, conn, xmin, xmax = block(shape, nx, ny, nz, jitter)
    t = ["MESH", " NODES", "  PARAMETERS, 3, 3, 1", "  DISPLACEMENT DOF, 1, 2, 3", "  COORDINATES"]
    t += ["   %d, %.17g, %.17g, %.17g" % (n + 1, p[0], p[1], p[2]) for n, p in enumerate(X)]
    t += ["  END COORDINATES", " END NODES", " ELEMENTS, USER", "  PARAMETERS, %d, 0, 1001" % conn.shape[1], "  PROPERTIES",
          "   %.17g" % E, "   %.17g" % nu, "  END PROPERTIES", "  CONNECTIVITY"]
    for e, c in enumerate(conn):
        ids = [str(e + 1)] + [str(n) for n in c]
        line = "   " + ", ".join(ids)
        if len(line) > 95:                                         # the reference reads lines of at most 100 characters
            raise ValueError("connectivity line too long for the input format")
        t.append(line)
    t += ["  END CONNECTIVITY", " END ELEMENTS", "END MESH", "BOUNDARY CONDITIONS"]

    def nodeset(name, nodes):
        out = [" NODESET, " + name]
        for k in range(0, len(nodes), 12):
            out.append("  " + ", ".join(str(n) for n in nodes[k:k + 12]))
        return out + [" END NODESET"]

    t += nodeset("xmin", xmin) + nodeset("xmax", xmax)
    t += [" DEGREES OF FREEDOM", "  xmin, 1, VALUE, 0.d0", "  xmin, 2, VALUE, 0.d0", "  xmin, 3, VALUE, 0.d0", " END DEGREES OF FREEDOM",
          " FORCES"] + ["  xmax, %d, VALUE, %.17g" % (d + 1, f) for d, f in enumerate(force) if f != 0.0] + [" END FORCES"]
    t += ["END BOUNDARY CONDITIONS", "STATIC STEP", " INITIAL TIME STEP, 1.d0", " MAX TIME STEP, 1.d0", " MIN TIME STEP, 0.001d0",
          " MAX NUMBER OF STEPS, 1", " STOP TIME, 1.d0", " SOLVER, %s, LINEAR" % solver, " CG TOLERANCE, %.17g" % cg_tolerance,
          " CG MAX ITERATIONS, %d" % cg_maxit, "END STATIC STEP", "STOP", ""]
    return "\n".join(t)


def tied_block_text(nx, ny, nz, cut=None, jitter=0.15, E=100.0, nu=0.3, seed=11, solver="DIRECT, NONLINEAR, 1.d-10, 6",
                    tie_face_dof=None, force=(0.0, 0.0, -0.01), steps=1, cg_tolerance=1e-26, cg_maxit=20000, pull=None,
                    element="1001", props=None):
    """hex8 block as literal `.in` text with the CONSTRAINTS block of the reference's input format
    (read_input_file.f90:1429-1541).
    cut = i: the node plane x-index i is DUPLICATED — elements right of it use the copies — and the two sides are tied
             in all three DOFs by CONSTRAIN NODE PAIRS over two node sets: the solution is that of the uncut block.
    tie_face_dof = d: TIE NODES — DOF d of every node of the face x = max (but the first) follows the first node of that face.
    pull = v: DOF 1 of that master node is prescribed to v (HISTORY ramp) instead of the nodal forces.
    Returns (text, info) with info = {"orig": nodes of the uncut numbering, "copy": {orig node: its copy}, "master": node}."""
    rng = np.random.default_rng(seed)
    nid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i + 1
    X = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                p = np.array([i, j, k], float)
                if 0 < i < nx and 0 < j < ny and 0 < k < nz:
                    p += jitter * rng.uniform(-1, 1, 3)
                X[nid(i, j, k) - 1] = p
    n_orig = X.shape[0]
    copy = {}
    if cut is not None:
        extra = []
        for k in range(nz + 1):
            for j in range(ny + 1):
                copy[nid(cut, j, k)] = n_orig + len(extra) + 1
                extra.append(X[nid(cut, j, k) - 1])
        X = np.vstack([X, np.array(extra)])
    conn = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)]
                if cut is not None and i >= cut:
                    c = [copy.get(n, n) for n in c]
                conn.append(c)
    xmin = [nid(0, j, k) for k in range(nz + 1) for j in range(ny + 1)]
    xmax = [nid(nx, j, k) for k in range(nz + 1) for j in range(ny + 1)]
    t = ["MESH", " NODES", "  PARAMETERS, 3, 3, 1", "  DISPLACEMENT DOF, 1, 2, 3", "  COORDINATES"]
    t += ["   %d, %.17g, %.17g, %.17g" % (n + 1, p[0], p[1], p[2]) for n, p in enumerate(X)]
    t += ["  END COORDINATES", " END NODES", " ELEMENTS, USER", "  PARAMETERS, 8, 0, %s" % element, "  PROPERTIES"]
    t += ["   %.17g" % v for v in (props if props is not None else (E, nu))] + ["  END PROPERTIES", "  CONNECTIVITY"]
    t += ["   " + ", ".join([str(e + 1)] + [str(n) for n in c]) for e, c in enumerate(conn)]
    t += ["  END CONNECTIVITY", " END ELEMENTS", "END MESH", "BOUNDARY CONDITIONS", " HISTORY, ramp", "  0.d0, 0.d0",
          "  %d.d0, 1.d0" % steps, " END HISTORY"]

    def nodeset(name, nodes):
        out = [" NODESET, " + name]
        for q in range(0, len(nodes), 12):
            out.append("  " + ", ".join(str(n) for n in nodes[q:q + 12]))
        return out + [" END NODESET"]

    t += nodeset("xmin", xmin) + nodeset("xmax", xmax)
    master = xmax[0]
    if cut is not None:
        t += nodeset("cutleft", sorted(copy)) + nodeset("cutright", [copy[n] for n in sorted(copy)])
    if tie_face_dof is not None:
        t += nodeset("followers", xmax[1:])
    t += [" DEGREES OF FREEDOM", "  xmin, 1, VALUE, 0.d0", "  xmin, 2, VALUE, 0.d0", "  xmin, 3, VALUE, 0.d0"]
    if pull is not None:
        t += ["  %d, 1, HISTORY, ramp" % master]
    t += [" END DEGREES OF FREEDOM"]
    if pull is None:
        t += [" FORCES"] + ["  xmax, %d, VALUE, %.17g" % (d + 1, f) for d, f in enumerate(force) if f != 0.0] + [" END FORCES"]
    t += [" CONSTRAINTS"]
    if cut is not None:
        t += ["  CONSTRAIN NODE PAIRS, %d, cutleft, %d, cutright, %d" % (len(copy), d, d) for d in (1, 2, 3)]
    if tie_face_dof is not None:
        t += ["  TIE NODES, %d, followers, %d, %d, %d" % (len(xmax) - 1, tie_face_dof, master, tie_face_dof)]
    t += [" END CONSTRAINTS", "END BOUNDARY CONDITIONS", "STATIC STEP", " INITIAL TIME STEP, 1.d0", " MAX TIME STEP, 1.d0",
          " MIN TIME STEP, 0.001d0", " MAX NUMBER OF STEPS, %d" % steps, " STOP TIME, %d.d0" % steps, " SOLVER, %s" % solver,
          " CG TOLERANCE, %.17g" % cg_tolerance, " CG MAX ITERATIONS, %d" % cg_maxit, "END STATIC STEP", "STOP", ""]
    text = "\n".join(t)
    if pull is not None:
        text = text.replace("  %d.d0, 1.d0" % steps, "  %d.d0, %.17g" % (steps, pull))
    return text, {"n_orig": n_orig, "copy": copy, "master": master, "xmax": xmax}


def order_with_constraints(arrays):
    """Equation order for the oracle's skyline of a model with tie constraints: nodes in natural order, every constraint
    equation right before the LAST of its two nodes (the reference lets its bandwidth minimiser place them among the nodes;
    appended after all DOFs they would follow the zero pivots of a body that is held by its ties only)."""
    N = int(arrays["n_nodes"][0])
    last = np.maximum(arrays["con_node1"], arrays["con_node2"]) - 1
    before = {}
    for k in range(last.size):
        before.setdefault(int(last[k]), []).append(N + k)
    out = []
    for n in range(N):
        out += before.get(n, [])
        out.append(n)
    return np.array(out, np.int32)
