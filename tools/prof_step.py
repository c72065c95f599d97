"""One short static step for profilers (ncu): assemble -> BC -> `iters` PCG iterations -> assemble -> BC on a jittered block.
  python tools/prof_step.py N method jitter iters       method 0 assembled, 1 matrix-free
Keeps the number of kernel launches small; prints the library's phase times (meaningless under a profiler)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input

n, method, jitter, iters = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
m = Model.from_text(synthetic_block_input(n, jitter=jitter))
fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
g = GpuSystem(m)
g.set_operator_mode(method)
z = np.zeros(g.ndof)
g.upload_state(z, z, None, fx, fd, fv)
for rep in range(2):
    g.zero_increment_dev()
    g.assemble_dev()
    g.apply_boundaryconditions_dev()
    g.solve_dev(method, 1e-20, iters)          # stops at the iteration cap (fail flag set: expected here)
    st = g.stats()
    print("rep %d: n=%d method=%d congruent=%d  assemble %.3f ms  bc %.3f ms  solve %.3f ms for %d iterations (%.1f us each)"
          % (rep, n, method, st.mf_congruent, st.assemble_ms, st.bc_ms, st.solve_ms, st.last_cg_iterations,
             1e3 * st.solve_ms / max(st.last_cg_iterations, 1)))
g.close()
