set -x
timeout 300 python -m pytest tests/test_mpnn_gpu.py -m gpu -x -q -k edge_mlp 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); e=d['mpnn']['edge_mlp']; print(e['tcgen05'], e['fp32_pipe']['ms'], e['forward_backward']['ms'], e['tensor'])"
