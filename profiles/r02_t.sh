set -x
T=r02_t
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py tests/test_core_step_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
python bench.py --steps 20 --warmup 5 --no-mpnn --no-ppo --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$T.json')); e=d['e2e']; print(d['ms_per_step'], e['value']/1e9, e['windows_ms'], e['host_enqueue_ms'], e['delta_tt_edge_form_reproduced_on_host'])"
python bench.py --steps 100 --warmup 5 --no-mpnn --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); e=d['e2e']; print(d['ms_per_step'], e['value']/1e9, e['windows_ms'], e['host_enqueue_ms'])"
