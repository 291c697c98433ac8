set -x
T=r02_q
python -m pytest tests/test_mpnn_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -4 gpurun_out/pytest_$T.log
python profiles/value_train_prof.py 2>&1 | grep -E "k_value|Self CUDA time"
