mkdir -p gpurun_out
timeout 600 python tools/multi_bench.py 8 > gpurun_out/r02x_multi8.json 2> gpurun_out/r02x_multi8.err; echo "multi8 rc=$?"; cat gpurun_out/r02x_multi8.json; tail -3 gpurun_out/r02x_multi8.err
timeout 300 python tools/multi_bench.py 2 > gpurun_out/r02x_multi2.json 2> gpurun_out/r02x_multi2.err; echo "multi2 rc=$?"; cat gpurun_out/r02x_multi2.json
timeout 300 ./tests/c/multi_check 24 24 64 8 0,1,2,3,4,5,6,7 0
timeout 300 ./tests/c/multi_check 16 16 32 4 0,1,2,3 1
