set -x
T=r02_x
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant mb10 "-DTARL_SELECT_MINBLOCKS=10"
build_variant mb8 "-DTARL_SELECT_MINBLOCKS=8"
for v in base mb10 mb8; do
  lib=/tmp/$v/libtarl.so; [ $v = base ] && lib=tarl_simulator_b200/libtarl_b200.so
  TARL_B200_LIB=$lib python bench.py --steps 20 --warmup 5 --no-mpnn --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v', d['ms_per_step'], d['roofline']['kernels_ms'], d['roofline']['frac'], d['roofline']['step']['frac'], d['e2e']['value']/1e9)"
done
for a in 300 500 1000; do
TARL_AHEAD_SELECT=$a python bench.py --steps 20 --warmup 5 --no-mpnn --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('ahead $a', d['ms_per_step'], d['roofline']['kernels_ms'], d['roofline']['frac'])"
done
