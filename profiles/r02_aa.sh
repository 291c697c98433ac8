set -x
timeout 300 python -m pytest tests/test_mpnn_gpu.py -m gpu -x -q -k "edge_mlp" 2>&1 | tail -15
