"""Cost of the tie-constraint layer: an n^3 block cut at its middle node plane and tied node to node (3 (n+1)^2 constraints),
next to the uncut block.   python tools/tie_times.py N"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshes
from en234_fea_b200 import GpuSystem, Model

n = int(sys.argv[1])
out = {}
for name, cut in (("uncut", None), ("cut + tied", n // 2)):
    text, info = meshes.tied_block_text(n, n, n, cut=cut, solver="CONJUGATE GRADIENT, LINEAR", cg_tolerance=1e-24, cg_maxit=50000)
    m = Model.from_text(text)
    g = GpuSystem(m)
    g.run_static(method=0, check_linear_cg=False)          # warm-up (graph capture)
    t0 = time.perf_counter()
    res = g.run_static(method=0, check_linear_cg=False)
    wall = time.perf_counter() - t0
    st = g.stats()
    out[name] = res["dof_total"][:3 * info["n_orig"]]
    print("n=%d %-11s constraints %6d: PCG iterations %5d, solve %.3f ms (%.2f us / iteration), assemble %.3f ms, bc %.3f ms, step wall %.1f ms"
          % (n, name, m.ints("con_node1").size, res["cg_iterations"][0], st.solve_ms, 1e3 * st.solve_ms / max(1, res["cg_iterations"][0]),
             st.assemble_ms, st.bc_ms, 1e3 * wall), flush=True)
    g.close()
print("max |u(cut + tied) - u(uncut)| / max |u| = %.2e" % (np.abs(out["cut + tied"] - out["uncut"]).max() / np.abs(out["uncut"]).max()))
