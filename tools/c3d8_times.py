"""Assembly time of the built-in C3D8 B-bar + elastic UMAT element (k_element_c3d8 + the ordered row gather) on a synthetic
block, next to the user element 1001 (k_element_blocks) on the same mesh.   python tools/c3d8_times.py N [reps]"""
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
base = Model.from_text(synthetic_block_input(n, jitter=0.2)).text()
head = base[:base.index(" ELEMENTS, USER")]
conn = base[base.index("  CONNECTIVITY"):]
c3d8 = (head + " MATERIAL, elastic_umat\n  STATE VARIABLES, 0\n  PROPERTIES\n   100\n   0.3\n  END PROPERTIES\n END MATERIAL\n"
        " ELEMENTS, INTERNAL\n  TYPE, C3D8\n  PROPERTIES, elastic_umat\n" + conn)
c3d8 = c3d8.replace(" SOLVER, CONJUGATE GRADIENT, LINEAR", " SOLVER, DIRECT, NONLINEAR, 1.d-8, 15, UNSYMMETRIC")
c3d8 = re.sub(r" DISTRIBUTED LOADS.*?END DISTRIBUTED LOADS\n", " FORCES\n  xmax, 3, VALUE, -1.d-4\n END FORCES\n", c3d8, flags=re.S)
for name, text in (("1001 (k_element_blocks)", base), ("C3D8 + UMAT (k_element_c3d8)", c3d8)):
    m = Model.from_text(text)
    g = GpuSystem(m)
    z = np.zeros(g.ndof)
    fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
    g.upload_state(z, z, None, fx, fd, fv)
    t = []
    for _ in range(reps):
        g.assemble_dev()
        t.append(g.stats().assemble_ms)
    print("n=%d %-30s assemble ms %s" % (n, name, ["%.3f" % x for x in t]))
    g.close()
