set -x
T=r02_k
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py tests/test_core_step_gpu.py tests/test_sim_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
TARL_AHEAD_SELECT=1500 TARL_AHEAD_RESPOND=2500 python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py -m gpu -x -q > gpurun_out/pytest_${T}_ahead.log 2>&1; tail -3 gpurun_out/pytest_${T}_ahead.log
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant nokeep "-DTARL_L2_KEEP=0"
build_variant rb12 "-DTARL_RESPOND_MINBLOCKS=12"
: > gpurun_out/tune_$T.log
TARL_TUNE=nokeep TARL_B200_LIB=/tmp/nokeep/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=keep python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=rb12 TARL_B200_LIB=/tmp/rb12/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
for pol in 00 10 12 02 11; do TARL_TUNE=pol$pol TARL_L2_POLICY=$pol python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1; done
TARL_TUNE=rb12 TARL_B200_LIB=/tmp/rb12/libtarl.so TARL_AHEAD_SELECT=700 TARL_AHEAD_RESPOND=1200 python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
for pol in 00 01 10 11 21; do TARL_TUNE=pol$pol TARL_L2_POLICY=$pol python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1; done
for a in "700 1200"; do
  set -- $a
  TARL_TUNE=keep TARL_AHEAD_SELECT=$1 TARL_AHEAD_RESPOND=$2 python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
done
TARL_TUNE=keep python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=keep TARL_AHEAD_SELECT=700 TARL_AHEAD_RESPOND=1200 python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=nokeep TARL_B200_LIB=/tmp/nokeep/libtarl.so python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
grep -v Warn gpurun_out/tune_$T.log
python bench.py --steps 20 --no-mpnn --no-ppo --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_k.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernels_ms'], d['roofline']['frac'], d['roofline']['step']['frac'])
PY
