set -x
T=r02_p
python -m pytest tests/test_ppo_device_gpu.py tests/test_sim_gpu.py tests/test_runner_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
for R in 128 1024; do
python profiles/rollout_timeline.py $R 2>&1 | grep -E "^R |kernel time"
TARL_NO_ROLLOUT_OVERLAP=1 python profiles/rollout_timeline.py $R 2>&1 | grep -E "^R |kernel time"
done
