"""Device-timed phases of the static step on one GPU (scratch tool for kernel work; not part of bench.py).
  python tools/phase_times.py N [element] [reps]   -> assemble / BC milliseconds (CUDA events inside the library)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input

n = int(sys.argv[1])
element = sys.argv[2] if len(sys.argv) > 2 else "1001"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kw = dict(jitter=0.2, element=element)
if element in ("1002", "U2"):
    kw["props"] = [1.0, 200.0]
m = Model.from_text(synthetic_block_input(n, **kw))
fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
g = GpuSystem(m)
z = np.zeros(g.ndof)
g.upload_state(z, z, None, fx, fd, fv)
asm, bc = [], []
for _ in range(reps):
    g.assemble_dev()
    asm.append(g.stats().assemble_ms)
    g.apply_boundaryconditions_dev()
    bc.append(g.stats().bc_ms)
E_ = m.n_elements
nnz = 9 * g.nnz_blocks()
alg = 32 * E_ + 72 * m.n_nodes + 64 * E_ + 8 * nnz + 16 * g.ndof
print("n=%d element=%s EN234_ASM=%s  assemble ms %s  (algorithmic %.3f GB -> %.0f GB/s at best)  bc ms %s" % (
    n, element, os.environ.get("EN234_ASM", ""), ["%.3f" % t for t in asm], alg / 1e9, alg / 1e6 / min(asm),
    ["%.3f" % t for t in bc]))
g.close()
