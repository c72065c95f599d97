"""Wall time of the state-print projection (en234gpu_project_fields through the host API) on an n^3 jittered block."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = Model.from_text(synthetic_block_input(n, jitter=0.2))
g = GpuSystem(m)
rng = np.random.default_rng(0)
u = 0.01 * rng.uniform(-1, 1, g.ndof)
names = ["S11", "S22", "S33", "S12", "S13", "S23", "SMISES"]
for lumped in (False, True):
    g.project_fields(names, u, lumped=lumped)
    k0 = g.stats().kernel_launches
    t = time.perf_counter()
    f = g.project_fields(names, u, lumped=lumped)
    dt = time.perf_counter() - t
    print("project_fields n=%d^3 nodes=%d fields=7 lumped=%s: %.1f ms wall (h2d of u + d2h of fields included), %d kernel launches"
          % (n, m.n_nodes, lumped, 1e3 * dt, g.stats().kernel_launches - k0), flush=True)
g.close()
