set -x
T=r02_am
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -4 gpurun_out/pytest_$T.log
for V in "" "TARL_NO_UNIFORM_WEIGHTS=1"; do
  echo "== $V"; env $V python profiles/tune_step.py 7 20 2>&1 | tail -1
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mpnn --no-ppo 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['per_kernel'], d['roofline']['step']['frac'], d['e2e']['value'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mpnn --no-ppo --workload grid100 --replicas 1024 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('grid100x1024', d['ms_per_step'], d['value'], d['roofline']['per_kernel'], d['roofline']['step']['frac'])"
