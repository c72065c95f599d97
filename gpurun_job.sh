python -m pytest tests -m gpu -x -q 2>&1 | tail -4
./en234_fea_b200/csrc/en234_run tests/golden/linear_elastic_3d_uel.in 2>&1 | tail -8
