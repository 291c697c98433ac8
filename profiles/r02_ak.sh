set -x
timeout 600 python -m pytest tests/test_ppo_device_gpu.py tests/test_runner_gpu.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_ak.json 2> gpurun_out/bench_r02_ak.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_ak.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value']); print(d['ppo'])"
