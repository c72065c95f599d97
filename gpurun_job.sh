mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep "mgpu_check\|rror\|Traceback" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-1500 | tail -20
echo "bench rc=$?"
