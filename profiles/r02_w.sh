set -x
T=r02_w
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py tests/test_core_step_gpu.py tests/test_sim_gpu.py tests/test_ppo_device_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant rb16 "-DTARL_RESPOND_MINBLOCKS=16"
build_variant rb10 "-DTARL_RESPOND_MINBLOCKS=10"
: > gpurun_out/tune_$T.log
for v in base rb16 rb10 base; do
  lib=/tmp/$v/libtarl.so; [ $v = base ] && lib=tarl_simulator_b200/libtarl_b200.so
  TARL_TUNE=$v TARL_B200_LIB=$lib python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
done
TARL_TUNE=base python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=base python profiles/tune_step.py 3 20 grid100 1 >> gpurun_out/tune_$T.log 2>&1
grep -v Warn gpurun_out/tune_$T.log
python bench.py --steps 20 --warmup 5 --no-mpnn --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['kernels_ms'], d['roofline']['frac'], d['roofline']['step']['frac'], d['e2e']['value']/1e9)"
python profiles/rollout_timeline.py 128 2>&1 | grep -E "^R |kernel time"
python profiles/rollout_timeline.py 1024 2>&1 | grep -E "^R |kernel time"
