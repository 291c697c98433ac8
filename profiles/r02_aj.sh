set -x
timeout 600 python -m pytest tests/test_ppo_device_gpu.py tests/test_runner_gpu.py -m gpu -x -q 2>&1 | tail -3
python profiles/rollout_timeline.py 128 2>&1 | grep -E "^R |checksum"
python bench.py --steps 20 --warmup 5 --no-mpnn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('no-mpnn ppo:', {k:d['ppo'][k] for k in ('rollout_ms','iteration_ms','update_ms')})"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('full    ppo:', {k:d['ppo'][k] for k in ('rollout_ms','iteration_ms','update_ms')})"
