set -x
T=r02_e
python -m pytest tests/test_ppo_device_gpu.py tests/test_runner_gpu.py tests/test_mpnn_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -25 gpurun_out/pytest_$T.log
python bench.py --steps 20 --no-mpnn > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -5 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_e.json'))
print(json.dumps(d.get('ppo'), indent=1))
print(d['value'], d['e2e']['value'], d['roofline']['kernels_ms'])
PY
