import sys
import numpy as np
sys.path.insert(0, ".")
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input
lumped = sys.argv[1] == "lumped"
m = Model.from_text(synthetic_block_input(4, jitter=0.2))
g = GpuSystem(m)
u = 0.01 * np.random.default_rng(0).uniform(-1, 1, g.ndof)
f = g.project_fields(["S11", "S22", "S33", "SMISES"], u, lumped=lumped)
print("lumped", lumped, np.abs(f).max(0), flush=True)
g.close()
