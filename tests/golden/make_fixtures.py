"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists):

  linear_elastic_3d_uel.in   re-emission (by the product's own writer, 17 significant digits, no comments) of the
  holeplate_3d_uel.in.gz     mesh/BC/step data parsed from input_files/Abaqus_uel_linear_elastic_3d.in and
                             input_files/Abaqus_uel_holeplate_3d.in — data fixtures, the GPU box has no /root/reference
  stretch_rotate_3d_umat.in  the same for input_files/Abaqus_umat_stretch_rotate.in (SUBROUTINE-type prescribed DOFs)
  linear_elastic_3d_umat.in  the same for input_files/Abaqus_umat_linear_elastic_3d.in (two C3D8 + UMAT elements)
  holeplate_3d_umat.in.gz    the same for input_files/Abaqus_umat_holeplate_3d.in (C3D8 + UMAT, the golden-log case)
  *_oracle_output.npz        OUTPUTS OF THE REPO'S ORACLE (not of the reference, which cannot run here): displacements, reaction forces and Newton history of the CPU oracle (skyline direct
                             solver path) on those inputs
"""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/EN234_FEA/input_files"

from en234_fea_b200 import Model  # noqa: E402
import oracle_binding as ob  # noqa: E402

for src, dst in [("Abaqus_uel_linear_elastic_3d.in", "linear_elastic_3d_uel.in"),
                 ("Abaqus_uel_holeplate_3d.in", "holeplate_3d_uel.in.gz")]:
    m = Model.read(os.path.join(REF, src))
    text = m.text()
    path = os.path.join(HERE, dst)
    if dst.endswith(".gz"):
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(text.encode())
    else:
        open(path, "w").write(text)
    m2 = Model.from_text(text)
    a, b = m.arrays(), m2.arrays()
    for k in a:
        if k in ("cg_tolerance", "max_cg_iterations", "checkstiffness_flag"):
            continue
        assert np.array_equal(a[k], b[k]), k
    om = ob.OracleModel(a)
    res = om.run_static()
    np.savez_compressed(os.path.join(HERE, dst.split(".")[0] + "_oracle_output.npz"), dof_total=res["dof_total"],
                        rforce=res["rforce"], newton=res["newton"], profile=np.array(res["profile"]), lu=res["lu"])
    print(dst, "nodes", m.n_nodes, "elements", m.n_elements, "steps", res["n_steps"], "newton rows", len(res["newton"]))
    print(res["newton"])

# golden-log case (C3D8 + UMAT): the model arrays as parsed from input_files/Abaqus_umat_holeplate_3d.in; the expected
# Newton history is the reference's own Output_Files/Abaqus_umat_holeplate_3d.out (hard-coded in test_oracle_cpu.py)
mu = Model.read(os.path.join(REF, "Abaqus_umat_holeplate_3d.in"))
np.savez_compressed(os.path.join(HERE, "holeplate_3d_umat_model.npz"), **mu.arrays())
print("holeplate_3d_umat_model.npz", mu.n_nodes, mu.n_elements)
# the same model re-emitted as .in text by the product's writer, for the GPU run of the golden-log case
text = mu.text()
with gzip.GzipFile(os.path.join(HERE, "holeplate_3d_umat.in.gz"), "wb", mtime=0) as f:
    f.write(text.encode())
a, b = mu.arrays(), Model.from_text(text).arrays()
for k in a:
    if k in ("cg_tolerance", "max_cg_iterations", "checkstiffness_flag"):
        continue
    assert np.array_equal(a[k], b[k]), k
print("holeplate_3d_umat.in.gz round trip ok")
# the two-element C3D8 + UMAT input (uniaxial tension by a NORMAL distributed load, 3 steps, NONLINEAR 1e-9)
m2 = Model.read(os.path.join(REF, "Abaqus_umat_linear_elastic_3d.in"))
open(os.path.join(HERE, "linear_elastic_3d_umat.in"), "w").write(m2.text())
# the same bar stretched 1 % and then rotated rigidly by 90 degrees through SUBROUTINE-type prescribed DOFs
# (user_prescribeddof): input_files/Abaqus_umat_stretch_rotate.in
m3 = Model.read(os.path.join(REF, "Abaqus_umat_stretch_rotate.in"))
open(os.path.join(HERE, "stretch_rotate_3d_umat.in"), "w").write(m3.text())
# explicit dynamics: the two VUEL inputs (user_codes/src/abaqus_vuel.for), EXPLICIT DYNAMIC STEP blocks with their PRINT STATE
for src, dst in [("Abaqus_vuel_linear_elastic_3d.in", "linear_elastic_3d_vuel.in"), ("Abaqus_vuel_holeplate_3d.in", "holeplate_3d_vuel.in.gz")]:
    mv = Model.read(os.path.join(REF, src))
    text = mv.text()
    if dst.endswith(".gz"):
        with gzip.GzipFile(os.path.join(HERE, dst), "wb", mtime=0) as f:
            f.write(text.encode())
    else:
        open(os.path.join(HERE, dst), "w").write(text)
    a, b = mv.arrays(), Model.from_text(text).arrays()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert mv.state_print() == Model.from_text(text).state_print()
    print(dst, "nodes", mv.n_nodes, "elements", mv.n_elements, "dt", a["dynamic_timestep"][0], "steps", a["no_dynamic_steps"][0])
