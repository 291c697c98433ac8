set -x
timeout 600 python -m pytest tests/test_ppo_device_gpu.py tests/test_runner_gpu.py -m gpu -x -q 2>&1 | tail -5
python profiles/ppo_update_prof.py 128 2>&1 | grep -v Warn | tail -32
python profiles/rollout_timeline.py 128 2>&1 | grep -v Warn | tail -24
