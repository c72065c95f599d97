mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_handle" > gpurun_out/r02w_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/r02w_multi_tests.log
timeout 600 $T 29511 tests/mgpu_check.py > gpurun_out/r02w_mgpu_all.log 2>&1; echo "mgpu all rc=$?"; grep mgpu_check gpurun_out/r02w_mgpu_all.log | cut -c1-250
timeout 600 $T 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-secondary 2> gpurun_out/r02w_n2.err | grep '^{' > gpurun_out/r02w_n2.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02w_n2.json"))
print("N=2 value %.0f e2e %.0f iter ms %.4f spmv ms %.4f parity %s" % (d["value"], d["e2e"]["value"], d["roofline"]["cg_iteration"]["ms"], d["roofline"]["ms_per_launch"], d["parity"]["ok"]), d["clocks"].get("sm_mhz"))
PY
