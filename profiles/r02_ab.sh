set -x
T=r02_ab
timeout 300 python -m pytest tests/test_mpnn_gpu.py -m gpu -x -q -k "edge_mlp" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-ppo --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$T.json')); e=d['mpnn']['edge_mlp']; print(e['tcgen05']['ms'], e['fp32_pipe']['ms'], e['forward_backward']['ms'], e['tensor'])"
