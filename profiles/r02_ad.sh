set -x
T=r02_ad
timeout 600 python -m pytest tests/test_mpnn_gpu.py tests/test_runner_gpu.py -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --no-ppo --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$T.json')); m=d['mpnn']
for k in ('policy_distribution','value_net','value_net_train_mode'): print(k, m[k]['ms_per_iter'], m[k]['roofline']['frac'])"
