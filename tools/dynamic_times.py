"""Device time per explicit dynamic step (en234gpu_dynamic_steps, CUDA events) on an n^3 jittered block of VU1 elements."""
import sys

sys.path.insert(0, ".")
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 320
m = Model.from_text(synthetic_block_input(n, jitter=0.2, element="U1", props=[100.0, 0.3, 1.0], load="stretch", steps=4, dt=0.5,
                                          dynamic=(1e-4 * 64 / n, steps)))
g = GpuSystem(m)
g.run_dynamic(32)                      # set-up + warm-up (graph instantiation)
for rep in range(3):
    fail, ms = g.dynamic_steps(1e-4 * 64 / n, steps)
    E, nd = m.n_elements, 3 * m.n_nodes
    # per step: element kernel reads 8 nodes x 6 doubles per element (cached) and writes 24 doubles; the vector kernels move
    # du, f (write) ; v, a, du, f, mass, ut, Re (read) ; a, v, ut, rforce (write)
    bytes_step = E * (32 + 192 + 192) + nd * 8 * 13
    print("dynamic n=%d^3 elements=%d: %d steps in %.2f ms = %.1f us/step, %.2f G element-steps/s, ~%.0f GB/s algorithmic, fail=%d"
          % (n, E, steps, ms, 1e3 * ms / steps, E * steps / ms / 1e6, bytes_step * steps / ms / 1e6, fail), flush=True)
g.close()
