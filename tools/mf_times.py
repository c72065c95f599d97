"""Operator-application timings on one GPU: assembled block-CSR SpMV vs matrix-free (scratch tool).
  python tools/mf_times.py N [jitter]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input

n = int(sys.argv[1])
jitter = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
m = Model.from_text(synthetic_block_input(n, jitter=jitter))
fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
g = GpuSystem(m)
z = np.zeros(g.ndof)
g.upload_state(z, z, None, fx, fd, fv)
out = {}
for method in (0, 1):
    g.set_operator_mode(method)
    g.assemble_dev()
    asm = []
    for _ in range(3):
        g.assemble_dev()
        asm.append(g.stats().assemble_ms)
    g.apply_boundaryconditions_dev()
    out[method] = (min(asm), g.stats().bc_ms, g.bench_operator(method, 5, 50))
print("n=%d jitter=%g  assembled: assemble %.3f ms bc %.3f ms apply %.4f ms | matrix-free: residual assembly %.3f ms bc %.3f ms apply %.4f ms"
      % ((n, jitter) + out[0] + out[1]))
g.close()
