mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_v2.json 2> gpurun_out/bench_n1_v2.err; echo rc=$?; tail -2 gpurun_out/bench_n1_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_v2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline']['cg_iteration'], d['roofline']['phases_ms'], d['clocks'])
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/bench_c3_v2.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3_v2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline']['cg_iteration'], d['roofline']['phases_ms'], d['clocks'])
PY
