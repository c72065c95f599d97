mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo rc=$?
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_spmv|k_assemble_rows|k_cg_update|k_cg_pupdate|k_apply_bc' -c 8 -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo rc=$?
