set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 64 128; do python tools/phase_times.py $n; done
python tools/phase_times.py 96 1002
ncu --set full --import-source on --clock-control none -k regex:"k_element_blocks|k_gather_rows" -c 2 -o gpurun_out/asm2k_b -f python tools/phase_times.py 64 1001 1 > gpurun_out/ncu_asm.log 2>&1
