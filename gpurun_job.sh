for rep in 1 2; do for occ in 4 5; do for n in 64 128; do EN234_SPMV_OCC=$occ python ab_test.py $n; done; done; done
