set -x
T=r02_f
python -m pytest tests/test_mpnn_gpu.py tests/test_ppo_device_gpu.py tests/test_runner_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -25 gpurun_out/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -6 gpurun_out/smoke_$T.log
python profiles/ppo_update_prof.py 128 2>&1 | grep "^R " 
python profiles/ppo_update_prof.py 1024 2>&1 | grep "^R "
python bench.py --steps 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -5 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_f.json'))
print(json.dumps(d.get('ppo'), indent=1))
print(json.dumps(d['mpnn']['value_mlp'], indent=1))
print(d['value'], d['e2e']['value'], d['roofline']['kernels_ms'])
PY
