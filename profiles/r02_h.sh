set -x
T=r02_h
python -m pytest tests/test_mpnn_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -15 gpurun_out/pytest_$T.log
bash profiles/r02_g.sh 2>&1 | grep " R "
python bench.py --steps 20 --no-ppo --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_h.json'))
m=d['mpnn']
for k in ('policy_distribution','value_net','value_net_train_mode'):
    print(k, m[k]['ms_per_iter'], m[k].get('roofline',{}).get('frac'))
PY
python profiles/value_train_prof.py 2>&1 | tail -25
