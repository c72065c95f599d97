python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'P'
import gzip, time, sys
sys.path.insert(0, '.')
from en234_fea_b200 import GpuSystem, Model
m = Model.from_text(gzip.open('tests/golden/holeplate_3d_umat.in.gz','rb').read().decode())
g = GpuSystem(m)
t = time.time(); res = g.run_static(method=0); dt = time.time() - t
print("holeplate C3D8 run: %.2f s, BiCGSTAB iterations per solve %s, phases ms %s" % (dt, res["cg_iterations"].tolist(), (g.stats().assemble_ms, g.stats().bc_ms, g.stats().solve_ms)))
P
