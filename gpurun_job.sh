set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1_e.log 2>&1; echo rc=$?
ncu --set full --clock-control none -c 16 -o gpurun_out/prof_r1e -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2_e.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:"k_stencil_nodes|k_mask_soa|k_element_blocks|k_mf_gather" -s 12 -c 6 -o gpurun_out/prof_r1e_mf -f python tools/mf_times.py 128 0.0 > gpurun_out/ncu3_e.log 2>&1; echo rc=$?
ls -la gpurun_out
