set -x
T=r02_ac
timeout 900 python -m pytest tests/test_sim_gpu.py tests/test_ppo_device_gpu.py tests/test_runner_gpu.py tests/test_metrics.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -12 gpurun_out/pytest_$T.log
python profiles/rollout_timeline.py 128 2>&1 | grep -E "^R |kernel time|k_insert"
python profiles/rollout_timeline.py 1024 2>&1 | grep -E "^R |kernel time|k_insert"
